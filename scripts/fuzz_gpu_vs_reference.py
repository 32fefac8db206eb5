"""Differential fuzzing of the CUDA path (through the C ABI, device-resident and host-layout calls)
against the translated reference (oracle/_ref/libbgc_ref.so) with the input / parameter generators
of scripts/fuzz_oracle_vs_reference.py.  Tolerances: tests/parity.py.  Rounds whose reference
output contains NaN / Inf are compared on their finite cells only and counted.

    python scripts/fuzz_gpu_vs_reference.py [rounds] [seed]          (on a B200 box)
    python scripts/fuzz_gpu_vs_reference.py --dry [rounds] [seed]    (CPU: the oracle stands in for the GPU,
                                                                      checks this script's own logic)
"""
import os
import sys
import threading

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, os.path.join(REPO, "scripts"))
import parity                               # noqa: E402
import ref_translated as rt                 # noqa: E402
import fuzz_oracle_vs_reference as gen      # noqa: E402

o = parity.oracle()
DRY = "--dry" in sys.argv
args = [a for a in sys.argv[1:] if a != "--dry"]


def finite_nerr(got, ref, floor=0.0):
    """max |got - ref| / max(max |ref|, floor) over the cells where the reference is finite"""
    m = np.isfinite(ref)
    if not m.any():
        return 0.0
    g, r = np.where(m, got, 0.0), np.where(m, ref, 0.0)
    if floor > 0.0 and np.isfinite(g).all():
        return float(np.max(np.abs(g - r)) / max(float(np.max(np.abs(r))), floor))
    return parity.nerr(g, r)


# Ill-conditioned by construction in the reference.  In the bottom cell OTHER_REMIN is
# min(.., flux - POC_sed - SED_DENITRIF*dz*denitrif_C_N) / dz (BGC_mod.F90:2539-2549) and both
# DENITRIF (:1574-1575) and O2_CONSUMPTION (:1786-1790, hence the O2 tendency) then form
# (.. - OTHER_REMIN)/denitrif_C_N - SED_DENITRIF resp. (.. - SED_DENITRIF*denitrif_C_N - OTHER_REMIN):
# whenever the second argument of that min() wins, +-SED_DENITRIF cancels and what remains is the
# rounding noise of the cancelled terms.  SED_DENITRIF contains 0.99**(O2 - NO3) (:2532-2534), which the
# fuzzer's NO3 of 10^4..10^5 mmol/m3 drives to 10^100..10^170; the reference's own result is then
# noise of relative size 2^-53 * SED_DENITRIF (or an exact 0 where its IEEE division happens to
# round back), and so is anybody's.  These three outputs are therefore measured against the
# magnitude of the terms that cancel, like the conservation residuals in tests/parity.py.
CANCELLING = ("diag_DENITRIF", "diag_O2_CONSUMPTION")
DENITRIF_C_N = 117.0 / 136.0


def one_round(seed, report):
    rng = np.random.default_rng(seed)
    host = parity.pkg.host
    parms = o.Parms() if DRY else host.Parms()
    if seed % 3:
        gen.perturb_parms(parms, rng)
    rp = rt.RefParms(parms).sync_from(parms)
    nL, nC = int(rng.choice([3, 17, 60])), int(rng.choice([1, 7, 256, 256, 257]))   # incl. the one-column calls of upstream MPAS
    cols, dms, mac = parity.make_bgc(nL, nC, parms, with_dms=True, with_macros=True, seed=1000 + seed)
    gen.perturb_inputs(cols, dms, mac, rng)
    alt = bool(rng.integers(0, 2))
    device_mode = bool(seed % 2)
    ref = cols.copy()
    rt.BGC_SourceSink(rp, ref, alt)
    if DRY:
        got = cols.copy()
        o.BGC_SourceSink(parms, got, alt)
    else:
        ctx = host.Context(nL, nC, device=0, parms=parms)
        got = parity.run_gpu_bgc(ctx, cols.copy(), device_mode=device_mode, alt_co2_use_eco=alt)
    worst, where = 0.0, ""
    act = ref.active_mask()
    sed = np.where(act & np.isfinite(ref.diag["diag_SedDenitrif"]), ref.diag["diag_SedDenitrif"], 0.0)
    cancel_scale = float(np.max(np.abs(sed / np.where(act, ref.cell_thickness, 1.0)), initial=0.0)) * DENITRIF_C_N
    for n in range(30):
        e = finite_nerr(got.BGC_tendencies[:, :, n], ref.BGC_tendencies[:, :, n],
                        floor=cancel_scale if n + 1 == parms.ind.o2_ind else 0.0)
        if e > worst:
            worst, where = e, "tendency %d" % (n + 1)
    for nm, a in ref.diag.items():
        if nm in parity.SOLVER_DIAGS or nm.startswith("diag_Jint"):
            continue
        e = finite_nerr(got.diag[nm], a, floor=cancel_scale if nm in CANCELLING else 0.0)
        if e > worst:
            worst, where = e, nm
    eph = finite_nerr(np.where(act, got.PH_PREV_3D, 0.0), np.where(act, ref.PH_PREV_3D, 0.0))
    dref, dgot, mref, mgot = dms.copy(), dms.copy(), mac.copy(), mac.copy()
    rt.DMS_SourceSink(rp, dref); rt.MACROS_SourceSink(rp, mref)
    if DRY:
        o.DMS_SourceSink(parms, dgot); o.MACROS_SourceSink(parms, mgot)
    else:
        host.DMS_SourceSink(ctx, dgot); host.MACROS_SourceSink(ctx, mgot)
        st = ctx.status()
        ctx.close()
    e = max(finite_nerr(dgot.DMS_tendencies, dref.DMS_tendencies), finite_nerr(mgot.MACROS_tendencies, mref.MACROS_tendencies))
    if e > worst:
        worst, where = e, "DMS/MACROS tendencies"
    report[seed] = dict(worst=worst, where=where, ph=eph, cancel_scale=cancel_scale, nonfinite=int((~np.isfinite(ref.BGC_tendencies)).sum()),
                        mode="device" if device_mode else "host", status=None if DRY else st)


def main():
    rounds = int(args[0]) if args else 100
    seed0 = int(args[1]) if len(args) > 1 else 0
    report = {}
    for s in range(seed0, seed0 + rounds):
        t = threading.Thread(target=one_round, args=(s, report))
        t.start(); t.join()
        if s not in report:
            print("seed", s, "crashed"); return 1
    bad = {s: r for s, r in report.items() if not (r["worst"] <= parity.TOL_TEND and r["ph"] <= parity.TOL_SOLVER)}
    print("%d rounds (%s), worst tendency/diagnostic error %.3e, worst pH error %.3e, %d rounds beyond tolerance"
          % (rounds, "dry: oracle as the implementation" if DRY else
             "CUDA path, %s flavour, seeds %d.." % (os.environ.get("BGC_B200_FLAVOUR", "prod"), seed0),
             max(r["worst"] for r in report.values()), max(r["ph"] for r in report.values()), len(bad)))
    for s, r in sorted(bad.items())[:20]:
        print("  seed %d (%s): %.3e at %s, pH %.3e, status %s" % (s, r["mode"], r["worst"], r["where"], r["ph"], r["status"]))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
