"""Differential fuzzing of the hand-written oracle against the translated reference
(oracle/_ref/libbgc_ref.so): random parameter tables, functional-group tables, switches, and inputs
pushed far outside the synthetic profiles (zeros, tiny and huge concentrations, anoxic and
supersaturated water, polar / tropical / hot columns, very shallow and ragged bathymetry, dark and
bright columns, huge dust and iron fluxes, ice fractions outside [0,1]), so that rarely taken
branches run.  Every output of every routine must agree bit for bit (NaN == NaN).

    python scripts/fuzz_oracle_vs_reference.py [rounds] [seed]
"""
import os
import sys
import threading

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import parity                    # noqa: E402
import ref_translated as rt      # noqa: E402

o = parity.oracle()
abi = parity.abi


def perturb_parms(po, rng):
    b = po.bgc
    for n in ("parm_Fe_bioavail", "parm_o2_min", "parm_o2_min_delta", "parm_kappa_nitrif", "parm_nitrif_par_lim",
              "parm_z_mort_0", "parm_z_mort2_0", "parm_labile_ratio", "parm_POMbury", "parm_BSIbury",
              "parm_fe_scavenge_rate0", "parm_f_prod_sp_CaCO3", "parm_POC_diss", "parm_SiO2_diss", "parm_CaCO3_diss"):
        setattr(b, n, getattr(b, n) * rng.uniform(0.3, 2.5))
    for i in range(4):
        b.parm_scalelen_vals[i] *= rng.uniform(0.5, 2.0)
    b.lrest_po4, b.lrest_no3, b.lrest_sio3 = (int(x) for x in rng.integers(0, 2, 3))
    for a in po.autotrophs:
        for n, t in a._fields_:
            if n.endswith("_ind") or n in ("Nfixer", "imp_calcifier", "exp_calcifier"):
                continue
            if n == "temp_function":
                a.temp_function = int(rng.integers(1, 3))
            else:
                setattr(a, n, getattr(a, n) * rng.uniform(0.5, 1.8))
        a.grazee_ind = int(rng.integers(1, 5))
        if rng.random() < 0.3:
            a.kSiO3 = float(rng.choice([0.0, 0.4]))
        a.temp_optN, a.temp_optS = a.temp_thresN - rng.uniform(2, 12), a.temp_thresS - rng.uniform(2, 12)
    for src in (po.dms, po.macros):
        for n, _ in src._fields_:
            setattr(src, n, getattr(src, n) * rng.uniform(0.5, 2.0))


def perturb_inputs(cols, dms, mac, rng):
    nL, nC = cols.nLevelsMax, cols.nColumnsMax
    scale = rng.choice([0.0, 1e-9, 1e-4, 0.1, 1.0, 1.0, 1.0, 5.0, 1e3], size=(nL, nC, 30))
    cols.BGC_tracers *= scale
    cols.BGC_tracers[rng.random((nL, nC, 30)) < 0.02] = -0.5           # the clamp
    cols.PotentialTemperature[...] = rng.choice([-1.9, 0.0, 4.0, 15.0, 29.0, 36.0], size=(nL, nC)) + rng.normal(0, 0.5, (nL, nC))
    cols.Salinity[...] = rng.choice([0.05, 5.0, 30.0, 35.0, 40.0], size=(nL, nC))
    cols.cell_latitude[...] = rng.uniform(-1.5, 1.5, nC)
    kmax = rng.integers(0, nL + 1, nC).astype(np.int32)
    kmax[rng.random(nC) < 0.3] = rng.integers(1, 4)                     # bottom above 100 m
    for c in (cols, dms, mac):
        c.number_of_active_levels[...] = kmax
    F = cols.forcing
    F["ShortWaveFlux_surface"][...] = rng.choice([0.0, 0.5, 50.0, 400.0, 1200.0], size=nC)
    F["dust_FLUX_IN"][...] = rng.choice([0.0, 1e-13, 1e-11, 1e-8], size=nC)
    F["FESEDFLUX"][...] = rng.choice([0.0, 2.3e-6, 1e-3], size=(nL, nC))
    F["iceFraction"][...] = rng.choice([-0.3, 0.0, 0.4, 1.0, 1.7], size=nC)
    F["windSpeedSquared10m"][...] = rng.choice([0.0, 1.0, 1e4, 2.25e6, 1e8], size=nC)
    F["surfacePressure"][...] = rng.uniform(0.9, 1.1, nC)
    F["atmCO2"][...] = rng.choice([180.0, 400.0, 1200.0], size=nC)
    F["SST"][...] = cols.PotentialTemperature[0]
    F["SSS"][...] = cols.Salinity[0]
    F["NUTR_RESTORE_RTAU"][...] = rng.uniform(0.0, 1e-6, (nL, nC))
    for nm, slot in (("NO3_CLIM", 2), ("PO4_CLIM", 1), ("SiO3_CLIM", 3)):
        F[nm][...] = np.abs(cols.BGC_tracers[:, :, slot - 1]) * rng.uniform(0.5, 1.5, (nL, nC))
    for nm in ("depositionFlux", "riverFlux", "seaIceFlux"):
        F[nm][...] = rng.normal(0.0, 1e-6, F[nm].shape)
    cols.PH_PREV_3D[...] = rng.choice([0.0, 0.0, 7.9, 8.2, 5.5, 9.5], size=(nL, nC))
    cols.PH_PREV_ALT_CO2_3D[...] = rng.choice([0.0, 8.0], size=(nL, nC))
    F["surface_pH"][...] = rng.choice([0.0, 8.1, 6.0], size=nC)
    cols.lcalc_O2_gas_flux, cols.lcalc_CO2_gas_flux = (int(x) for x in rng.integers(0, 2, 2))
    dms.DMS_tracers *= rng.choice([0.0, 1e-6, 1.0, 1.0, 100.0], size=dms.DMS_tracers.shape)
    mac.MACROS_tracers *= rng.choice([0.0, 1e-6, 1.0, 1.0, 100.0], size=mac.MACROS_tracers.shape)
    for k in ("ShortWaveFlux_surface", "iceFraction", "windSpeedSquared10m", "SST", "SSS", "surfacePressure"):
        dms.forcing[k][...] = F[k]


def differ(a, b):
    return not np.array_equal(a, b, equal_nan=True)


def one_round(seed, report):
    rng = np.random.default_rng(seed)
    po = o.Parms()
    if seed % 3:
        perturb_parms(po, rng)
    rp = rt.RefParms(po).sync_from(po)
    nL, nC = int(rng.choice([3, 17, 60])), 48
    cols, dms, mac = parity.make_bgc(nL, nC, po, with_dms=True, with_macros=True, seed=1000 + seed)
    perturb_inputs(cols, dms, mac, rng)
    alt = bool(rng.integers(0, 2))
    a, b = cols.copy(), cols.copy()
    da, db, ma, mb = dms.copy(), dms.copy(), mac.copy(), mac.copy()
    bad = []
    for p in range(2):
        o.BGC_SourceSink(po, a, alt); rt.BGC_SourceSink(rp, b, alt)
        if differ(a.BGC_tendencies, b.BGC_tendencies): bad.append("tend pass %d" % p)
        if differ(a.PH_PREV_3D, b.PH_PREV_3D) or differ(a.PH_PREV_ALT_CO2_3D, b.PH_PREV_ALT_CO2_3D): bad.append("pH pass %d" % p)
        bad += ["%s pass %d" % (n, p) for n in a.diag if differ(a.diag[n], b.diag[n])]
    o.BGC_SurfaceFluxes(po, a); rt.BGC_SurfaceFluxes(rp, b)
    bad += ["forcing " + n for n in a.forcing if differ(a.forcing[n], b.forcing[n])]
    bad += ["flux diag " + n for n in a.flux_diag if differ(a.flux_diag[n], b.flux_diag[n])]
    o.DMS_SourceSink(po, da); rt.DMS_SourceSink(rp, db); o.DMS_SurfaceFluxes(po, da); rt.DMS_SurfaceFluxes(rp, db)
    if differ(da.DMS_tendencies, db.DMS_tendencies): bad.append("DMS tend")
    bad += ["DMS " + n for n in da.diag if differ(da.diag[n], db.diag[n])]
    bad += ["DMS flux " + n for n in da.flux_diag if differ(da.flux_diag[n], db.flux_diag[n])]
    if differ(da.forcing["netFlux"], db.forcing["netFlux"]): bad.append("DMS netFlux")
    o.MACROS_SourceSink(po, ma); rt.MACROS_SourceSink(rp, mb)
    if differ(ma.MACROS_tendencies, mb.MACROS_tendencies): bad.append("MACROS tend")
    bad += ["MACROS " + n for n in ma.diag if differ(ma.diag[n], mb.diag[n])]
    report[seed] = (bad, int(np.isnan(b.BGC_tendencies).sum()), int(a.active_mask().sum()))


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    report = {}
    for s in range(seed0, seed0 + rounds):
        t = threading.Thread(target=one_round, args=(s, report))   # fresh module state per round
        t.start(); t.join()
        if s not in report:
            print("seed", s, "crashed"); return 1
    nbad = sum(1 for v in report.values() if v[0])
    cells = sum(v[2] for v in report.values())
    nan_rounds = sum(1 for v in report.values() if v[1])
    print("%d rounds, %d active cells, %d rounds with NaN outputs (agreeing), %d rounds with differences"
          % (rounds, cells, nan_rounds, nbad))
    for s, v in sorted(report.items()):
        if v[0]:
            print("  seed %d: %s" % (s, ", ".join(v[0][:8])))
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
