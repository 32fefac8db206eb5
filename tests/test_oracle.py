"""CPU tests that pin the ORACLE (oracle/, test infrastructure): the reference
ships no tests or golden vectors, so the restatement is anchored on the known
answers quoted in the reference's comments, on published check values of the
equilibrium constants, on an independent NumPy restatement of the carbonate
system, on the code's own conservation diagnostics, and on one unit test per
quirk of SURVEY.md section 8."""
import math
import os
import sys

import numpy as np
import pytest

import parity

pkg = parity.pkg
abi = pkg.abi
o = parity.oracle()
sys.path.insert(0, os.path.join(parity.REPO, "oracle"))
import co2calc_numpy as cn   # noqa: E402


@pytest.fixture(scope="module")
def po():
    return o.Parms()


def test_o2sat_check_value():
    # BGC_mod.F90:3028-3029: T = 10 degC, S = 35 permil -> 282.015 mmol/m^3
    assert abs(o.O2SAT(10.0, 35.0) - 282.015) < 5e-4


def test_dust_to_fe_check_value():
    # BGC_mod.F90:2491: 0.035 / 55.847 * 1e9 = 626712.0 nmol Fe / g dust
    assert abs(o.lib().oracle_dust_to_Fe() - 626712.0) < 1.0


def test_single_precision_literals(po):
    # BGC_parms.F90:480-482 under plain gfortran -O2: REAL(4) values widened (quirk Q6)
    assert po.bgc.epsC == float(np.float32(1.00e-8))
    assert po.bgc.epsTinv == float(np.float32(3.17e-8))
    assert po.bgc.epsC != 1.00e-8
    p8 = o.Parms(default_real_8=True)
    assert p8.bgc.epsC == 1.00e-8 and p8.bgc.epsTinv == 3.17e-8


def test_published_equilibrium_constants():
    """DOE (1994) / Dickson, Sabine & Christian (2007) check values at S = 35, t = 25 degC."""
    c = o.co3_coeffs([1], [0.0], [25.0], [35.0])
    ln = {k: math.log(v[0]) for k, v in c.items() if v[0] > 0}
    assert abs(ln["k0"] - (-3.5617)) < 2e-3
    assert abs(math.log10(c["k1"][0]) - (-5.8472)) < 1e-3
    assert abs(math.log10(c["k2"][0]) - (-8.9660)) < 1e-3
    assert abs(ln["kb"] - (-19.7964)) < 2e-3
    assert abs(ln["kw"] - (-30.434)) < 2e-3
    assert abs(ln["ks"] - (-2.30)) < 1e-2
    assert abs(ln["k1p"] - (-3.71)) < 1e-2
    assert abs(ln["k2p"] - (-13.727)) < 2e-3
    assert abs(ln["k3p"] - (-20.24)) < 1e-2
    assert abs(ln["ksi"] - (-21.61)) < 1e-2
    # total boron / sulfate at S = 35 (Uppstrom 1974; Morris & Riley 1966)
    assert abs(c["bt"][0] - 4.157e-4) < 2e-6
    assert abs(c["st"][0] - 0.02824) < 2e-5


def test_constants_match_numpy_restatement():
    rng = np.random.default_rng(1)
    n = 500
    T = rng.uniform(-1.8, 31, n); S = rng.uniform(30, 38, n); D = rng.uniform(0, 5500, n)
    for k in (1, 7):
        c = o.co3_coeffs(np.full(n, k, np.int32), D, T, S)
        K = cn.constants(T, S, D, deep=(k > 1))
        for nm in ("k1", "k2", "kb", "k1p", "k2p", "k3p", "ksi", "kw", "ks", "kf", "bt", "st", "ff"):
            np.testing.assert_allclose(c[nm], K[nm], rtol=1e-12, err_msg="%s k=%d" % (nm, k))


def test_pressure_correction_is_keyed_on_level_index():
    # quirk Q3 (co2calc.F90:480...): same depth, k = 1 vs k = 2 differ; quirk Q2: k1/k2 do not
    a = o.co3_coeffs([1], [3000.0], [2.0], [34.7])
    b = o.co3_coeffs([2], [3000.0], [2.0], [34.7])
    assert a["k1"][0] == b["k1"][0] and a["k2"][0] == b["k2"][0]
    for nm in ("kb", "kw", "ks", "kf", "k1p", "k2p", "k3p", "ksi"):
        assert b[nm][0] > a[nm][0] * 1.05, nm


def test_solver_against_independent_bisection():
    pts = pkg.synth_co2_points(2000)
    r = o.co2calc_points(pts)
    h, K, dic = cn.solve_h(pts["temp"], pts["salt"], pts["dic"], pts["ta"], pts["pt"], pts["sit"])
    # The reference solver stops as soon as |dx| < xacc = 1e-10 mol/kg (co2calc.F90:53,977), so
    # its H+ is only guaranteed to that ABSOLUTE accuracy ("co2star accurate to 3 significant
    # figures", co2calc.F90:46-51); the exact root must lie within xacc of it.
    h_ref = 10.0 ** (-r["ph"])
    assert np.max(np.abs(h_ref - h)) < 1.0e-10
    assert np.median(np.abs(h_ref - h) / h) < 1e-5          # ...and it is usually far better
    # speciation from the oracle's own H+ reproduces its outputs
    co2star = dic * h_ref ** 2 / (h_ref ** 2 + K["k1"] * h_ref + K["k1"] * K["k2"]) * 1e6 * 1.026
    np.testing.assert_allclose(r["co2star"], co2star, rtol=1e-9)
    pco2 = co2star / (1e6 * 1.026) / K["ff"] * 1e6
    np.testing.assert_allclose(r["pco2surf"], pco2, rtol=1e-9)


def test_solver_iteration_counts():
    # co2calc.F90:857-863: "iterates about 12 times" cold, "about 5" with a +-0.5 bracket
    n = 5000
    pts = pkg.synth_co2_points(n)
    pts["phlo"][:] = 6.0; pts["phhi"][:] = 9.0
    cold = o.co2calc_points(pts)
    per = cold["stats"]["talk_row_calls"] / n
    assert 9.0 < per < 16.0, per
    assert cold["stats"]["no_convergence"] == 0 and cold["stats"]["bracket_grow"] == 0
    pts["phlo"] = cold["ph"] - 0.5; pts["phhi"] = cold["ph"] + 0.5
    warm = o.co2calc_points(pts)
    per = warm["stats"]["talk_row_calls"] / n
    assert 3.0 < per < 8.0, per
    # two different brackets, one root: both answers are within xacc of it
    assert np.max(np.abs(10.0 ** (-warm["ph"]) - 10.0 ** (-cold["ph"]))) < 2.0e-10


def test_bracket_growth_and_floors():
    # ALK below the bracket's reach: pH < 6 forces the growth loop (co2calc.F90:920-938);
    # DIC/ALK <= floors exercise dic_min/alk_min (:57-59, :843-846)
    r = o.comp_CO3terms(3, 100.0, 10.0, 35.0, 2000.0, 0.0, 1.0, 10.0, 6.0, 9.0)
    assert r["bracket_grow"] >= 1 and r["pH"] < 6.0 and np.isfinite(r["pH"])
    r2 = o.comp_CO3terms(3, 100.0, 10.0, 35.0, 0.0, 0.0, 0.0, 0.0, 6.0, 9.0)
    r3 = o.comp_CO3terms(3, 100.0, 10.0, 35.0, 1.0, 1.0, 0.0, 0.0, 6.0, 9.0)
    assert r2["pH"] == r3["pH"]


def _run(po, nL=40, nC=96, **kw):
    cols, _, _ = parity.make_bgc(nL, nC, po, **kw)
    o.BGC_SourceSink(po, cols, True, nthreads=4)
    return cols


def test_conservation_integrals_vanish(po):
    cols = _run(po, ragged=True)
    for el in ("C", "N", "P", "Si"):
        j = cols.diag["diag_Jint_%stot" % el]
        scale = np.max(np.abs(cols.diag["diag_Jint_100m_%stot" % el]))
        assert scale > 0
        assert np.max(np.abs(j)) < 1e-9 * scale, el


def test_alt_co2_solve_uses_dic_not_dic_alt(po):
    # quirk Q1 (BGC_mod.F90:975): DIC_ALT_CO2 never reaches BGC_SourceSink's outputs
    a, _, _ = parity.make_bgc(30, 32, po)
    b = a.copy()
    b.BGC_tracers[:, :, po.ind.dic_alt_co2_ind - 1] *= 1.37
    o.BGC_SourceSink(po, a, True); o.BGC_SourceSink(po, b, True)
    assert np.array_equal(a.BGC_tendencies, b.BGC_tendencies)
    assert np.array_equal(a.PH_PREV_ALT_CO2_3D, b.PH_PREV_ALT_CO2_3D)
    assert np.array_equal(a.PH_PREV_ALT_CO2_3D, a.PH_PREV_3D)


def test_negative_tracers_are_clamped(po):
    a, _, _ = parity.make_bgc(30, 32, po, jitter=False)
    b = a.copy()
    sl = po.ind.nh4_ind - 1
    a.BGC_tracers[:, :, sl] = 0.0
    b.BGC_tracers[:, :, sl] = -0.25
    o.BGC_SourceSink(po, a, True); o.BGC_SourceSink(po, b, True)
    assert np.array_equal(a.BGC_tendencies, b.BGC_tendencies)


def test_zero_mask_zeroes_the_whole_group(po):
    # BGC_mod.F90:826-844: Chl == 0 zeroes C, Fe, (Si, CaCO3) of that group
    a, _, _ = parity.make_bgc(20, 16, po, jitter=False)
    b = a.copy()
    at = po.autotrophs[1]   # diatoms
    a.BGC_tracers[:, :, at.Chl_ind - 1] = 0.0
    for i in (at.Chl_ind, at.C_ind, at.Fe_ind, at.Si_ind):
        b.BGC_tracers[:, :, i - 1] = 0.0
    o.BGC_SourceSink(po, a, True); o.BGC_SourceSink(po, b, True)
    assert np.array_equal(a.BGC_tendencies, b.BGC_tendencies)
    assert np.all(a.BGC_tendencies[:, :, at.C_ind - 1] == 0.0)


def test_whole_array_zero_semantics(po):
    # quirk Q15: tendencies and diagnostics are zero outside active cells, PH_PREV untouched there,
    # the three never-touched diagnostics keep the caller's values
    cols, _, _ = parity.make_bgc(30, 64, po, ragged=True, nColumns=50)
    parity.poison_outputs(cols)
    cols.PH_PREV_3D[...] = 0.0; cols.PH_PREV_ALT_CO2_3D[...] = 0.0
    o.BGC_SourceSink(po, cols, True)
    m = cols.active_mask()
    assert (~m).any() and m.any()
    assert np.all(cols.BGC_tendencies[~m] == 0.0)
    assert np.all(cols.PH_PREV_3D[~m] == 0.0) and np.all(cols.PH_PREV_3D[m] > 0.0)
    for nm in abi.BGC_DIAG_K2:
        if nm in abi.BGC_DIAG_UNTOUCHED:
            assert np.all(cols.diag[nm] == 7.25), nm
        else:
            assert np.all(cols.diag[nm][~m] == 0.0), nm
    for nm in abi.BGC_DIAG_KA:
        assert np.all(cols.diag[nm][~m] == 0.0), nm
    dead = cols.number_of_active_levels < 1
    dead[50:] = True
    for nm in abi.BGC_DIAG_C1:
        assert np.all(cols.diag[nm][dead] == 0.0), nm


def test_photoC_NO3_TOT_zint_double_accumulation(po):
    # quirk Q11 (BGC_mod.F90:1844-1846): the running per-group integral is added every level
    cols = _run(po, nL=12, nC=8, jitter=False)
    nL = 12
    dz = cols.cell_thickness[:, 0]
    per = cols.diag["diag_photoC_NO3"][:, 0, :] * dz[:, None]
    running = np.cumsum(per, axis=0)
    want = running.sum()
    got = cols.diag["diag_photoC_NO3_TOT_zint"][0]
    assert want > 0 and abs(got - want) <= 1e-12 * want
    plain = per.sum()
    assert abs(got - plain) > 1e-3 * plain


def test_surface_fluxes_side_effects(po):
    cols, _, _ = parity.make_bgc(10, 40, po)
    cols.forcing["iceFraction"][:5] = 1.7
    cols.forcing["iceFraction"][5:10] = -0.3
    p2 = o.Parms()
    p2.bgc.parm_Fe_bioavail = 0.5
    dep = cols.forcing["depositionFlux"].copy()
    o.BGC_SurfaceFluxes(p2, cols)
    fe = p2.ind.fe_ind - 1
    assert np.all(cols.forcing["iceFraction"][:5] == 1.0) and np.all(cols.forcing["iceFraction"][5:10] == 0.0)
    np.testing.assert_array_equal(cols.forcing["depositionFlux"][:, fe], dep[:, fe] * 0.5)   # quirk Q14
    other = [i for i in range(30) if i != fe]
    np.testing.assert_array_equal(cols.forcing["depositionFlux"][:, other], dep[:, other])
    assert np.all(cols.forcing["surface_pH"] > 7.0) and np.all(cols.forcing["surface_pH"] < 9.0)
    assert np.all(cols.flux_diag["pistonVel_O2"][:5] == 0.0)   # full ice cover
    alk, nh4, no3 = (getattr(p2.ind, n) - 1 for n in ("alk_ind", "nh4_ind", "no3_ind"))
    f = cols.forcing
    base = f["depositionFlux"] + f["gasFlux"] + f["riverFlux"] + f["seaIceFlux"]
    np.testing.assert_allclose(f["netFlux"][:, alk], base[:, alk] + f["netFlux"][:, nh4] - f["netFlux"][:, no3],
                               rtol=0, atol=1e-24)


def test_dms_and_macros_basic(po):
    _, dms, mac = parity.make_bgc(25, 48, po, with_dms=True, with_macros=True, ragged=True)
    parity.poison_outputs(dms); parity.poison_outputs(mac)
    o.DMS_SourceSink(po, dms); o.MACROS_SourceSink(po, mac)
    k = np.arange(1, 26)[:, None]
    m = k <= dms.number_of_active_levels[None, :]
    live = {po.dms_ind.dms_ind - 1, po.dms_ind.dmsp_ind - 1}
    for n in range(14):
        if n not in live:
            assert np.all(dms.DMS_tendencies[:, :, n] == 0.0)
    assert np.all(dms.DMS_tendencies[~m] == 0.0)
    # DMS / MACROS diagnostics are NOT zeroed outside active cells (quirk Q15)
    assert np.all(dms.diag["diag_DMS_S_TOTAL"][~m] == 7.25)
    assert np.all(mac.diag["diag_PROT_S_TOTAL"][~m] == 7.25)
    assert np.all(mac.MACROS_tendencies[~m] == 0.0)
    np.testing.assert_array_equal(dms.diag["diag_DMS_S_DMSP"][m], dms.diag["diag_DMS_S_TOTAL"][m])
