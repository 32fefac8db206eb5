"""Parity checker: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded synthetic columns.  Shared by the `-m gpu` tests and by
__graft_entry__.smoke().

Tolerances (SURVEY.md section 8(d) / BASELINE.json north_star):
  * tendencies and solver-independent diagnostics: per array,
        max|gpu - ref| / max|ref|  <=  1e-10
    (normalised by the array's own magnitude: pointwise relative error is
    ill-defined where tendencies cancel to ~0);
  * H+ of the carbonate solve: both sides stop when |dx| < xacc = 1e-10 mol/kg
    (co2calc.F90:53), so pH may legitimately differ by ~4e-3 in the worst case;
    in practice both follow the same Newton trajectory and agree to ~1e-12.
    The test bound for pH / CO3 / HCO3 / H2CO3 / co2star / pCO2 is 1e-8.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
abi = pkg.abi

TOL_TEND = 1e-10
TOL_SOLVER = 1e-8

# diagnostics whose value depends on the converged H+ of the carbonate solve
SOLVER_DIAGS = {"diag_CO3", "diag_HCO3", "diag_H2CO3", "diag_pH_3D", "diag_CO3_ALT_CO2",
                "diag_HCO3_ALT_CO2", "diag_H2CO3_ALT_CO2", "diag_pH_3D_ALT_CO2",
                "diag_zsatcalc", "diag_zsatarag"}
SOLVER_FLUX = {"co2star", "dco2star", "pco2surf", "dpco2", "co2star_alt_co2", "dco2star_alt_co2",
               "pco2surf_alt_co2", "dpco2_alt_co2"}


def oracle():
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle as o   # noqa: E402  (test infrastructure only)
    return o


def nerr(got, ref):
    """max|got-ref| / max|ref| (0 when both are identically zero)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if got.shape != ref.shape:
        raise AssertionError("shape mismatch %s vs %s" % (got.shape, ref.shape))
    bad = ~np.isfinite(got)
    if bad.any() and not (~np.isfinite(ref)).any():
        return float("inf")
    d = np.max(np.abs(got - ref)) if got.size else 0.0
    m = np.max(np.abs(ref)) if ref.size else 0.0
    if m == 0.0:
        return 0.0 if d == 0.0 else float("inf")
    return float(d / m)


def make_bgc(nL, nC, parms, *, nColumns=None, seed=None, ragged=False, jitter=True, column0=0,
             with_dms=False, with_macros=False, nlev_active=None):
    cols = pkg.BgcColumns(nL, nC, nColumns)
    dms = pkg.DmsColumns(nL, nC, nColumns) if with_dms else None
    mac = pkg.MacrosColumns(nL, nC, nColumns) if with_macros else None
    kw = {}
    if seed is not None:
        kw["seed"] = seed
    pkg.synth_fill(cols, dms, mac, bgc_ind=parms.ind, dms_ind=parms.dms_ind if with_dms else None,
                   macros_ind=parms.macros_ind if with_macros else None, ragged=ragged, jitter=jitter,
                   column0=column0, nlev_active=nlev_active, **kw)
    return cols, dms, mac


def poison_outputs(cols, value=7.25):
    """Fill every output with a sentinel so that 'never written' shows up."""
    for name in ("BGC_tendencies", "DMS_tendencies", "MACROS_tendencies"):
        if hasattr(cols, name):
            getattr(cols, name)[...] = value
    for a in getattr(cols, "diag", {}).values():
        a[...] = value
    for a in getattr(cols, "flux_diag", {}).values():
        a[...] = value


def compare_bgc_source_sink(ref, got, tol=TOL_TEND, tol_solver=TOL_SOLVER, diagnostics=True):
    """ref/got: BgcColumns after BGC_SourceSink.  Returns {field: error}; raises on failure."""
    errs = {}
    for n in range(abi.BGC_TRACER_CNT):
        errs["tend[%d]" % (n + 1)] = nerr(got.BGC_tendencies[:, :, n], ref.BGC_tendencies[:, :, n])
    errs["PH_PREV_3D"] = nerr(got.PH_PREV_3D, ref.PH_PREV_3D)
    errs["PH_PREV_ALT_CO2_3D"] = nerr(got.PH_PREV_ALT_CO2_3D, ref.PH_PREV_ALT_CO2_3D)
    if diagnostics:
        for nm, a in ref.diag.items():
            errs[nm] = nerr(got.diag[nm], a)
        # The conservation integrals are residuals of cancelling terms (BGC_mod.F90:1875-1938)
        # and are ~0 by construction - the full-column ones always, the upper-100 m ones when
        # the whole column lies above 100 m.  Measure their difference against the scale of
        # the terms being cancelled (column integral of |tendency| * dz of the element's main
        # inorganic pool), not against their own round-off noise.
        ind = pkg.host.Parms().ind
        dzm = np.where(ref.active_mask(), ref.cell_thickness, 0.0)
        for el, slot in (("C", ind.dic_ind), ("N", ind.no3_ind), ("P", ind.po4_ind), ("Si", ind.sio3_ind)):
            scale = np.max(np.sum(np.abs(ref.BGC_tendencies[:, :, slot - 1]) * dzm, axis=0))
            for nm in ("diag_Jint_%stot" % el, "diag_Jint_100m_%stot" % el):
                m = max(scale, np.max(np.abs(ref.diag[nm])))
                d = np.max(np.abs(got.diag[nm] - ref.diag[nm]))
                errs[nm] = float(d / m) if m > 0 else (0.0 if d == 0 else float("inf"))
    fails = []
    for k, e in errs.items():
        lim = tol_solver if (k in SOLVER_DIAGS or k.startswith("PH_PREV")) else tol
        if not (e <= lim):
            fails.append("%s: %.3e > %.1e" % (k, e, lim))
    if fails:
        raise AssertionError("BGC_SourceSink parity failed:\n  " + "\n  ".join(fails))
    return errs


def compare_fields(ref_dict, got_dict, tol, what, solver_keys=(), tol_solver=TOL_SOLVER, mask=None):
    errs, fails = {}, []
    for nm, a in ref_dict.items():
        g = got_dict[nm]
        if mask is not None and a.shape == mask.shape:
            errs[nm] = nerr(g[mask], a[mask])
        else:
            errs[nm] = nerr(g, a)
        lim = tol_solver if nm in solver_keys else tol
        if not (errs[nm] <= lim):
            fails.append("%s: %.3e > %.1e" % (nm, errs[nm], lim))
    if fails:
        raise AssertionError("%s parity failed:\n  %s" % (what, "\n  ".join(fails)))
    return errs


def run_gpu_bgc(ctx, cols, *, device_mode, alt_co2_use_eco=True, diagnostics=True, surface=False):
    """Run BGC_SourceSink (and optionally BGC_SurfaceFluxes) on the GPU; returns a
    host BgcColumns holding the results."""
    host = pkg.host
    out = cols.copy()
    if device_mode:
        dev = host.DeviceBgcColumns(cols.nLevelsMax, cols.nColumnsMax, cols.nColumns,
                                    device="cuda:%d" % ctx.device).load(out)
        host.BGC_SourceSink(ctx, dev, alt_co2_use_eco, diagnostics)
        if surface:
            host.BGC_SurfaceFluxes(ctx, dev)
        ctx.synchronize()
        dev.store(out)
    else:
        host.BGC_SourceSink(ctx, out, alt_co2_use_eco, diagnostics)
        if surface:
            host.BGC_SurfaceFluxes(ctx, out)
    return out


def reference_check(po, cols, oracle_result):
    """When oracle/_ref holds the translated reference (oracle/ref_translated.py), the oracle's
    answer for this block must equal the reference's bit for bit.  Returns a short verdict."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    try:
        import ref_translated as rt   # test infrastructure only
    except Exception as exc:          # noqa: BLE001
        return "unavailable (%s)" % exc
    if not rt.available():
        return "unavailable (oracle/_ref not shipped)"
    r = cols.copy()
    rt.BGC_SourceSink(rt.RefParms(po), r, True)
    assert np.array_equal(r.BGC_tendencies, oracle_result.BGC_tendencies), "oracle != reference (tendencies)"
    assert np.array_equal(r.PH_PREV_3D, oracle_result.PH_PREV_3D), "oracle != reference (pH)"
    for n, a in r.diag.items():
        assert np.array_equal(a, oracle_result.diag[n]), "oracle != reference (%s)" % n
    return "bit-identical"


def smoke(pkg_=None):
    """One small invocation of the hot path on cuda:0, checked against the oracle."""
    o = oracle()
    host = pkg.host
    parms = host.Parms()
    po = o.Parms()
    nL, nC = 24, 320
    cols, dms, mac = make_bgc(nL, nC, parms, ragged=True, with_dms=True, with_macros=True)
    ref = cols.copy()
    o.BGC_SourceSink(po, ref, True, nthreads=o.max_threads())
    pinned = reference_check(po, cols, ref)
    ctx = host.Context(nL, nC, device=0, parms=parms)
    got = run_gpu_bgc(ctx, cols, device_mode=True)
    errs = compare_bgc_source_sink(ref, got)
    worst = max(errs.values())

    # second (warm-bracket) pass
    ref2 = ref.copy(); got2_in = got.copy()
    o.BGC_SourceSink(po, ref2, True, nthreads=o.max_threads())
    got2 = run_gpu_bgc(ctx, got2_in, device_mode=True)
    worst = max(worst, max(compare_bgc_source_sink(ref2, got2).values()))

    # host-layout (Fortran) path
    got3 = run_gpu_bgc(ctx, cols, device_mode=False)
    worst = max(worst, max(compare_bgc_source_sink(ref, got3).values()))

    # co2calc_1point batch
    pts = pkg.synth_co2_points(4096)
    r = o.co2calc_points(pts, nthreads=o.max_threads())
    g = host.co2calc_points(ctx, pts)
    for k in ("ph", "co2star", "dco2star", "pco2surf", "dpco2"):
        e = nerr(g[k], r[k])
        assert e <= TOL_SOLVER, "co2calc_points %s: %.3e" % (k, e)
        worst = max(worst, e)

    # DMS / MACROS
    dref = dms.copy(); o.DMS_SourceSink(po, dref, nthreads=o.max_threads())
    dgot = dms.copy(); host.DMS_SourceSink(ctx, dgot)
    e = nerr(dgot.DMS_tendencies, dref.DMS_tendencies)
    assert e <= TOL_TEND, "DMS tendencies %.3e" % e
    mref = mac.copy(); o.MACROS_SourceSink(po, mref, nthreads=o.max_threads())
    mgot = mac.copy(); host.MACROS_SourceSink(ctx, mgot)
    e = nerr(mgot.MACROS_tendencies, mref.MACROS_tendencies)
    assert e <= TOL_TEND, "MACROS tendencies %.3e" % e

    st = ctx.status()
    assert st["no_bracket"] == 0 and st["no_convergence"] == 0, st
    print("smoke: worst normalised error %.3e; status %s; oracle vs translated reference on the same block: %s"
          % (worst, st, pinned))
    ctx.close()
