"""Pins the ORACLE against the reference itself.

oracle/_ref/libbgc_ref.so is the unmodified reference Fortran (/root/reference/*.F90)
machine-translated to C by oracle/f90c.py and compiled by gcc with the code generation
`gfortran -O2` uses on x86-64 (test infrastructure; see oracle/ref_translated.py).  Every test
below feeds the same inputs to the translated reference and to the hand-written oracle
(oracle/libbgc_oracle.so) and demands BIT-IDENTICAL results: the oracle restates the reference in
its evaluation order with the same libm, so anything short of equality is a restatement error.

The library is built wherever the reference sources exist and travels with the snapshot
otherwise; with neither, the tests are skipped (and the committed golden vectors, which were
produced by this library, still pin the oracle: tests/test_golden.py).
"""
import os
import sys
import threading

import numpy as np
import pytest

import parity

pkg = parity.pkg
abi = pkg.abi
o = parity.oracle()
sys.path.insert(0, os.path.join(parity.REPO, "oracle"))
import ref_translated as rt   # noqa: E402  (test infrastructure only)

if not rt.available() and rt.can_build():
    rt.build()
pytestmark = pytest.mark.skipif(not rt.available(), reason="oracle/_ref/libbgc_ref.so not built "
                                "(needs the reference sources once)")


@pytest.fixture(scope="module")
def po():
    return o.Parms()


@pytest.fixture(scope="module")
def rp(po):
    return rt.RefParms(po)


def same(a, b, what, mask=None):
    a, b = np.asarray(a), np.asarray(b)
    if mask is not None and a.shape[:mask.ndim] == mask.shape:
        a, b = a[mask], b[mask]
    if not np.array_equal(a, b, equal_nan=True):
        bad = a != b
        raise AssertionError("%s: %d of %d values differ, max |d| = %.3e (max |ref| = %.3e)" % (
            what, int(bad.sum()), a.size, float(np.nanmax(np.abs(a - b))), float(np.nanmax(np.abs(b)))))


def same_bgc(ref, got, what):
    same(got.BGC_tendencies, ref.BGC_tendencies, what + " tendencies")
    same(got.PH_PREV_3D, ref.PH_PREV_3D, what + " PH_PREV_3D")
    same(got.PH_PREV_ALT_CO2_3D, ref.PH_PREV_ALT_CO2_3D, what + " PH_PREV_ALT_CO2_3D")
    for n in ref.diag:
        same(got.diag[n], ref.diag[n], what + " " + n)


# ------------------------------------------------------------------ tables and constants
def test_parameter_tables(po, rp):
    """BGC_parms_init / DMS_parms_init / MACROS_parms_init (BGC_parms.F90:497-699,
    DMS_parms.F90:203-241, MACROS_parms.F90:143-162): every run-time tunable."""
    m = rt.meta()
    checked = 0
    for src, mod in ((po.bgc, "bgc_parms"), (po.dms, "dms_parms"), (po.macros, "macros_parms")):
        for n, _ in src._fields_:
            cn = "%s__%s" % (mod, n.lower())
            v = getattr(src, n)
            v = list(v) if hasattr(v, "__len__") else v
            if cn in m["vars"]:
                r = rt.var(cn)
                r = list(r) if hasattr(r, "__len__") else r.value
            elif cn in m["consts"]:
                r = rt.const(cn)
            else:
                assert n in ("lrest_po4", "lrest_no3", "lrest_sio3", "reserved", "T0_Kelvin_BGC") or \
                    n.startswith("lrest"), n
                continue
            assert v == r, (n, v, r)
            checked += 1
    assert checked >= 15 + 2 + 6 + 30 + 10 - 1


def test_functional_group_table_and_init(po, rp):
    """autotroph_type defaults (BGC_parms.F90:555-697) and BGC_init's index wiring and names
    (BGC_mod.F90:184-333)."""
    for i in range(4):
        for n, _ in abi.BgcAutotroph._fields_:
            assert getattr(po.autotrophs[i], n) == getattr(rp.autotrophs[i], n.lower()), (i, n)
    assert rp.autotrophs[0].sname.decode().rstrip() == "sp"
    assert rp.autotrophs[3].lname.decode().rstrip().lower().startswith("phaeo")
    names = [rp.name("ind", "short_name", i) for i in range(30)]
    assert names[:16] == ["PO4", "NO3", "SiO3", "NH4", "Fe", "O2", "DIC", "DIC_ALT_CO2", "ALK", "DOC",
                          "DON", "DOFe", "DOP", "DOPr", "DONr", "zooC"]
    assert sorted(names[16:]) == sorted(["spC", "spChl", "spFe", "spCaCO3", "diatC", "diatChl", "diatFe",
                                         "diatSi", "phaeoC", "phaeoChl", "phaeoFe", "diazC", "diazChl",
                                         "diazFe"])
    assert rp.name("ind", "units", 0) == "mmol/m^3"


def test_product_parameter_init_matches_the_reference(rp):
    """The C-ABI's own bgc_parms_init / bgc_init / dms_parms_init / macros_parms_init (host C in
    the product library, no GPU needed) against the reference's module variables and tables."""
    hp = pkg.host.Parms()
    m = rt.meta()
    for src, mod in ((hp.bgc, "bgc_parms"), (hp.dms, "dms_parms"), (hp.macros, "macros_parms")):
        for n, _ in src._fields_:
            cn = "%s__%s" % (mod, n.lower())
            v = getattr(src, n)
            v = list(v) if hasattr(v, "__len__") else v
            if cn in m["vars"]:
                r = rt.var(cn)
                assert v == (list(r) if hasattr(r, "__len__") else r.value), n
            elif cn in m["consts"]:
                assert v == rt.const(cn), n
    for i in range(4):
        for n, _ in abi.BgcAutotroph._fields_:
            assert getattr(hp.autotrophs[i], n) == getattr(rp.autotrophs[i], n.lower()), (i, n)
    for n, _ in abi.BgcIndices._fields_:
        assert getattr(hp.ind, n) == getattr(rp.ind, n.lower()), n


def test_named_constants_quirks():
    # Q6: single-precision literals widened (BGC_parms.F90:373, :480-486)
    assert rt.const("bgc_parms__epsc") == float(np.float32(1.00e-8)) != 1.00e-8
    assert rt.const("bgc_parms__epstinv") == float(np.float32(3.17e-8))
    assert rt.const("bgc_parms__epsnondim") == float(np.float32(1.00e-6))
    # Q5: integer(BGC_r8) "constants" of BGC_mod (BGC_mod.F90:120-125): p5 is 0
    assert rt.meta()["consts"]["bgc_mod__p5"] == "i8" and rt.const("bgc_mod__p5") == 0
    assert rt.const("bgc_mod__c10") == 10 and rt.const("bgc_mod__c1") == 1
    # co2calc.F90:53-59
    assert rt.const("co2calc__xacc") == 1e-10
    assert rt.const("co2calc__dic_min") == 0.1 / 35.0 * 1944.0
    assert rt.const("bgc_mod__bgc_tracer_cnt") == 30 and rt.const("dms_mod__dms_tracer_cnt") == 14
    assert rt.const("macros_mod__macros_tracer_cnt") == 8


def test_scalar_functions(rp):
    # BGC_mod.F90:3028-3029: the reference's own check value
    r, _ = rt.call("bgc_mod__o2sat_singlevalue", 10.0, 35.0)
    assert abs(r - 282.015) < 5e-4
    rng = np.random.default_rng(3)
    L = o.lib()
    for _ in range(300):
        t, s = rng.uniform(-2.0, 35.0), rng.uniform(0.0, 41.0)
        assert rt.call("bgc_mod__o2sat_singlevalue", t, s)[0] == o.O2SAT(t, s)
        assert rt.call("bgc_mod__schmidt_o2_singlevalue", t)[0] == L.oracle_SCHMIDT_O2_singleValue(t)
        assert rt.call("bgc_mod__schmidt_co2_singlevalue", t)[0] == L.oracle_SCHMIDT_CO2_singleValue(t)
        assert rt.call("dms_mod__schmidt_dms_singlevalue", t)[0] == L.oracle_SCHMIDT_DMS_singleValue(t)


# ------------------------------------------------------------------ carbonate system
def test_equilibrium_constants_and_saturation(rp):
    """comp_co3_coeffs (co2calc.F90:320-777) through its SAVE variables, comp_co3_sat_vals
    (:1096-1238); level 1 and deeper levels (Q2, Q3)."""
    rng = np.random.default_rng(5)
    n = 200
    T, S, D = rng.uniform(-1.8, 31, n), rng.uniform(0.05, 40, n), rng.uniform(0, 5500, n)
    for k in (1, 2, 37):
        c = o.co3_coeffs(np.full(n, k, np.int32), D, T, S)
        for i in range(n):
            _, b = rt.call("co2calc__comp_co3_coeffs", k, D[i], T[i], S[i], 0.0, 0.0, 0.0, 0.0, 1)
            got = dict(k0=b[4].value, k1=b[5].value, k2=b[6].value, ff=b[7].value)
            for nm in ("kw", "kb", "ks", "kf", "k1p", "k2p", "k3p", "ksi", "bt", "st"):
                got[nm] = rt.var("co2calc__" + nm)[0]
            for nm, v in got.items():
                assert v == c[nm][i], (k, i, nm, v, c[nm][i])
            assert rt.comp_co3_sat_vals(k, D[i], T[i], S[i]) == o.co3_sat_vals(k, D[i], T[i], S[i])


def test_comp_co3terms(rp):
    rng = np.random.default_rng(6)
    for i in range(400):
        k = int(rng.integers(1, 60))
        a = (k, rng.uniform(0, 5500), rng.uniform(-1.8, 31), rng.uniform(30, 38), rng.uniform(1800, 2400))
        a = a + (a[4] + rng.uniform(80, 420), rng.uniform(0, 3), rng.uniform(0, 150))
        lo, hi = (6.0, 9.0) if i % 2 else (lambda p: (p - 0.2, p + 0.2))(rng.uniform(7.6, 8.3))
        r = rt.comp_CO3terms(*a, lo, hi)
        g = o.comp_CO3terms(*a, lo, hi)
        for nm in ("pH", "H2CO3", "HCO3", "CO3"):
            assert r[nm] == g[nm], (i, nm, r[nm], g[nm])


def _oracle_1point(depth, temp, salt, dic, ta, pt, sit, phlo, phhi, xco2, atmpres, locmip=True):
    import ctypes as C
    lo, hi = C.c_double(phlo), C.c_double(phhi)
    out = [C.c_double() for _ in range(5)]
    o.lib().oracle_co2calc_1point(C.c_double(depth), C.c_int(int(locmip)), C.c_int(1), C.c_double(temp),
                                  C.c_double(salt), C.c_double(dic), C.c_double(ta), C.c_double(pt),
                                  C.c_double(sit), C.byref(lo), C.byref(hi), C.byref(out[0]), C.c_double(xco2),
                                  C.c_double(atmpres), *[C.byref(x) for x in out[1:]], None)
    return dict(zip(("ph", "co2star", "dco2star", "pco2surf", "dpco2"), (x.value for x in out)),
                phlo=lo.value, phhi=hi.value)


def test_branches_no_caller_takes(rp):
    """Statements of the reference that neither BGC_SourceSink nor BGC_SurfaceFluxes ever reach
    (found with gcov on the translated reference): the seawater-scale K1/K2 of
    comp_co3_coeffs (k1_k2_pH_tot = .false., co2calc.F90:466-469, :495-498) and the swapped-bracket
    branch of drtsafe_row (:940-947, entered when the caller's phlo > phhi)."""
    rng = np.random.default_rng(8)
    for i in range(200):
        a = (5.0, rng.uniform(-1.8, 31), rng.uniform(30, 38), rng.uniform(1800, 2300))
        a = a + (a[3] + rng.uniform(80, 420), rng.uniform(0, 3), rng.uniform(0, 150))
        lo, hi = (7.0, 9.0) if i % 2 == 0 else (9.0, 7.0)
        tail = (rng.uniform(280, 560), rng.uniform(0.95, 1.05))
        for locmip in (False, True):
            r = rt.co2calc_1point(*a, lo, hi, *tail, locmip_k1_k2_bug_fix=locmip)
            g = _oracle_1point(*a, lo, hi, *tail, locmip=locmip)
            for k in g:
                assert r[k] == g[k], (i, locmip, k, r[k], g[k])
    # the two pH scales differ, so the flag is not a no-op
    assert rt.co2calc_1point(*a, 7.0, 9.0, *tail, locmip_k1_k2_bug_fix=False)["ph"] != \
        rt.co2calc_1point(*a, 7.0, 9.0, *tail, locmip_k1_k2_bug_fix=True)["ph"]


def _ref_points(pts):
    n = len(pts["temp"])
    out = {k: np.zeros(n) for k in ("ph", "co2star", "dco2star", "pco2surf", "dpco2")}
    for i in range(n):
        r = rt.co2calc_1point(*[float(pts[k][i]) for k in
                                ("depth", "temp", "salt", "dic", "ta", "pt", "sit", "phlo", "phhi",
                                 "xco2", "atmpres")])
        for k in out:
            out[k][i] = r[k]
    return out


@pytest.mark.parametrize("warm", [False, True])
def test_co2calc_points(rp, warm):
    """BASELINE.json configs[1] (first 2048 of the 1M points), cold and warm brackets."""
    pts = pkg.synth_co2_points(2048)
    if warm:
        ph = o.co2calc_points(pts)["ph"]
        pts["phlo"], pts["phhi"] = ph - 0.2, ph + 0.2
    g, r = o.co2calc_points(pts), _ref_points(pts)
    for k in r:
        same(g[k], r[k], "co2calc_1point " + k)


def test_co2calc_points_extreme_inputs(rp):
    """Floors, brackets that miss the root (the growth loop of drtsafe_row), fresh water:
    the inputs of tests/test_gpu_parity.py::test_co2calc_points_extreme_inputs."""
    n = 1024
    rng = np.random.default_rng(5)
    pts = pkg.synth_co2_points(n)
    pts["temp"] = rng.choice([-1.9, 0.0, 12.0, 30.0, 35.0], size=n)
    pts["salt"] = rng.choice([0.02, 0.5, 5.0, 20.0, 35.0, 41.0], size=n)
    pts["dic"] = rng.choice([0.5, 3.0, 800.0, 2000.0, 2600.0], size=n)
    pts["ta"] = pts["dic"] * rng.uniform(0.9, 1.4, size=n)
    pts["pt"] = rng.choice([0.0, 0.5, 5.0], size=n)
    pts["sit"] = rng.choice([0.0, 20.0, 200.0], size=n)
    lo = rng.choice([3.0, 6.0, 7.0, 9.5], size=n)
    pts["phlo"], pts["phhi"] = lo, lo + rng.choice([0.4, 1.0, 2.0], size=n)
    g = o.co2calc_points(pts)
    ok = np.isfinite(g["ph"])     # a NaN residual never ends the reference's growth loop (Q9): skip those
    sel = {k: v[ok] for k, v in pts.items()}
    r = _ref_points(sel)
    for k in r:
        same(g[k][ok], r[k], "extreme co2calc_1point " + k)


# ------------------------------------------------------------------ the column routines
def test_single_column_config(po, rp):
    """BASELINE.json configs[0]: 1 column x 60 levels, cold then warm pass, then everything else."""
    cols, dms, mac = parity.make_bgc(60, 1, po, jitter=False, with_dms=True, with_macros=True)
    a, b = cols.copy(), cols.copy()
    for p in ("cold", "warm"):
        o.BGC_SourceSink(po, a, True)
        rt.BGC_SourceSink(rp, b, True)
        same_bgc(b, a, "single column " + p)
    o.BGC_SurfaceFluxes(po, a)
    rt.BGC_SurfaceFluxes(rp, b)
    for n in a.forcing:
        same(a.forcing[n], b.forcing[n], "forcing " + n)
    for n in a.flux_diag:
        same(a.flux_diag[n], b.flux_diag[n], "flux diag " + n)


@pytest.mark.parametrize("nL,nC,nCols,ragged", [(24, 96, 90, True), (60, 64, 64, False), (80, 40, 33, True)])
def test_bgc_source_sink_blocks(po, rp, nL, nC, nCols, ragged):
    cols, _, _ = parity.make_bgc(nL, nC, po, ragged=ragged, nColumns=nCols)
    parity.poison_outputs(cols)
    a, b = cols.copy(), cols.copy()
    for p in ("cold", "warm"):
        st = o.BGC_SourceSink(po, a, True, nthreads=2)
        rt.BGC_SourceSink(rp, b, True)
        same_bgc(b, a, "%dx%d %s" % (nL, nC, p))
        assert st["no_convergence"] == 0
    assert np.abs(a.BGC_tendencies).max() > 0
    # untouched members keep the sentinel in both (BGC_parms.F90 declares them, nothing writes them)
    for n in abi.BGC_DIAG_UNTOUCHED:
        assert np.all(b.diag[n] == 7.25) and np.all(a.diag[n] == 7.25)


def test_alt_co2_switch(po, rp):
    cols, _, _ = parity.make_bgc(30, 48, po, ragged=True)
    a, b = cols.copy(), cols.copy()
    o.BGC_SourceSink(po, a, False)
    rt.BGC_SourceSink(rp, b, False)
    same_bgc(b, a, "alt_co2_use_eco = .false.")


def _retune(parms):
    """the non-default tunables / group table of tests/test_gpu_parity.py::_retune"""
    b = parms.bgc
    b.parm_o2_min, b.parm_o2_min_delta = 6.0, 3.0
    b.parm_labile_ratio = 0.7
    b.parm_POMbury, b.parm_BSIbury = 1.3, 0.8
    b.parm_nitrif_par_lim = 2.5
    b.parm_kappa_nitrif *= 1.7
    b.parm_z_mort_0 *= 0.6
    b.parm_z_mort2_0 *= 1.4
    b.parm_fe_scavenge_rate0 *= 2.0
    b.parm_POC_diss, b.parm_SiO2_diss, b.parm_CaCO3_diss = 70.0e2, 300.0e2, 450.0e2
    b.parm_Fe_bioavail = 0.5
    for i in range(4):
        b.parm_scalelen_vals[i] *= (1.0 + 0.15 * i)
    b.lrest_no3 = b.lrest_po4 = b.lrest_sio3 = 1
    a = parms.autotrophs
    sp, diat, diaz, phaeo = (a[parms.ind.sp_ind - 1], a[parms.ind.diat_ind - 1], a[parms.ind.diaz_ind - 1],
                             a[parms.ind.phaeo_ind - 1])
    sp.temp_function, sp.temp_thresN, sp.temp_thresS, sp.temp_optN, sp.temp_optS = \
        abi.DEFINES["BGC_TFNC_QUASI_MMRT"], 27.0, 26.0, 18.0, 17.0
    phaeo.temp_function, phaeo.temp_thres = abi.DEFINES["BGC_TFNC_Q10"], 1.0
    diat.Qp = 0.0061
    sp.kSiO3 = 0.4
    diaz.graze_zoo, diaz.graze_poc, diaz.graze_doc = 0.25, 0.08, 0.10
    phaeo.grazee_ind = sp.grazee_ind
    diat.agg_rate_max, diat.agg_rate_min, diat.mort2 = 0.7, 0.03, 0.012
    for g in (sp, diat, diaz, phaeo):
        g.PCref *= 1.1
        g.alphaPI *= 0.9
    parms.dms.k_conv *= 1.3
    parms.dms.Stress_mult *= 0.8
    parms.macros.f_prot, parms.macros.k_poly_bac = 0.5, parms.macros.k_poly_bac * 2.0


def test_non_default_parameters_restoring_and_group_table():
    """Every branch keyed on a table value: QUASI_MMRT on one group and Q10 on Phaeocystis, the
    restoring terms (never switched on upstream, Q8), the remaining_P routing, a shared grazer,
    non-default DMS / MACROS rates; Fe_bioavail != 1 makes the in-place scaling of the forcing
    visible (Q14)."""
    nL, nC = 40, 96
    po2 = o.Parms()
    _retune(po2)
    res = {}

    def run():   # own thread: the retuned module variables must not leak into the other tests
        rp2 = rt.RefParms(po2).sync_from(po2)
        cols, dms, mac = parity.make_bgc(nL, nC, po2, ragged=True, seed=0x5EED, with_dms=True, with_macros=True)
        rng = np.random.default_rng(11)
        cols.forcing["NUTR_RESTORE_RTAU"][...] = rng.uniform(0.0, 1.0e-6, size=(nL, nC))
        tr = cols.BGC_tracers
        for nm, slot in (("NO3_CLIM", po2.ind.no3_ind), ("PO4_CLIM", po2.ind.po4_ind), ("SiO3_CLIM", po2.ind.sio3_ind)):
            cols.forcing[nm][...] = tr[:, :, slot - 1] * rng.uniform(0.8, 1.2, size=(nL, nC))
        cols.forcing["iceFraction"][:7] = [-0.2, 1.4, 0.3, 0.0, 1.0, 2.0, -1.0]
        a, b = cols.copy(), cols.copy()
        o.BGC_SourceSink(po2, a, True); rt.BGC_SourceSink(rp2, b, True)
        o.BGC_SurfaceFluxes(po2, a); rt.BGC_SurfaceFluxes(rp2, b)
        da, db = dms.copy(), dms.copy()
        o.DMS_SourceSink(po2, da); rt.DMS_SourceSink(rp2, db)
        o.DMS_SurfaceFluxes(po2, da); rt.DMS_SurfaceFluxes(rp2, db)
        ma, mb = mac.copy(), mac.copy()
        o.MACROS_SourceSink(po2, ma); rt.MACROS_SourceSink(rp2, mb)
        res.update(a=a, b=b, da=da, db=db, ma=ma, mb=mb)

    t = threading.Thread(target=run)
    t.start(); t.join()
    a, b = res["a"], res["b"]
    same_bgc(b, a, "retuned")
    assert np.abs(a.diag["diag_NO3_RESTORE"]).max() > 0
    for n in a.forcing:
        same(a.forcing[n], b.forcing[n], "retuned forcing " + n)
    for n in a.flux_diag:
        same(a.flux_diag[n], b.flux_diag[n], "retuned flux diag " + n)
    m = res["da"].active_mask() if hasattr(res["da"], "active_mask") else None
    same(res["da"].DMS_tendencies, res["db"].DMS_tendencies, "retuned DMS tendencies")
    for n in res["da"].diag:
        same(res["da"].diag[n], res["db"].diag[n], "retuned DMS " + n, m)
    for n in res["da"].forcing:
        same(res["da"].forcing[n], res["db"].forcing[n], "retuned DMS forcing " + n)
    same(res["ma"].MACROS_tendencies, res["mb"].MACROS_tendencies, "retuned MACROS tendencies")
    for n in res["ma"].diag:
        same(res["ma"].diag[n], res["mb"].diag[n], "retuned MACROS " + n, m)


def test_permuted_tracer_slots():
    """The host chooses the slots (BGC_indices_type, BGC_parms.F90:82-112)."""
    nL, nC = 30, 64
    perm = np.random.default_rng(7).permutation(30)
    po1 = o.Parms()
    po1.permute_tracers(perm)
    res = {}

    def run():
        rp1 = rt.RefParms(po1)
        cols, _, _ = parity.make_bgc(nL, nC, po1, ragged=True)
        a, b = cols.copy(), cols.copy()
        o.BGC_SourceSink(po1, a, True); rt.BGC_SourceSink(rp1, b, True)
        o.BGC_SurfaceFluxes(po1, a); rt.BGC_SurfaceFluxes(rp1, b)
        res.update(a=a, b=b)

    t = threading.Thread(target=run)
    t.start(); t.join()
    same_bgc(res["b"], res["a"], "permuted slots")
    same(res["a"].forcing["netFlux"], res["b"].forcing["netFlux"], "permuted netFlux")


@pytest.mark.parametrize("o2,co2", [(1, 1), (0, 1), (1, 0), (0, 0)])
def test_surface_fluxes(po, rp, o2, co2):
    """BGC_SurfaceFluxes (BGC_mod.F90:2706-2957) incl. the gas-flux switches and the in-place
    clamp of iceFraction."""
    cols, _, _ = parity.make_bgc(12, 200, po, nColumns=193)
    cols.lcalc_O2_gas_flux, cols.lcalc_CO2_gas_flux = o2, co2
    cols.forcing["iceFraction"][:7] = [-0.2, 1.4, 0.3, 0.0, 1.0, 2.0, -1.0]
    parity.poison_outputs(cols)
    a, b = cols.copy(), cols.copy()
    for p in ("cold", "warm"):
        o.BGC_SurfaceFluxes(po, a)
        rt.BGC_SurfaceFluxes(rp, b)
        for n in a.forcing:
            same(a.forcing[n], b.forcing[n], "%s forcing %s" % (p, n))
        for n in a.flux_diag:
            same(a.flux_diag[n], b.flux_diag[n], "%s flux diag %s" % (p, n))


def test_dms_and_macros(po, rp):
    """DMS_SourceSink / DMS_SurfaceFluxes / MACROS_SourceSink.  Their diagnostics are not zeroed
    by the reference (Q15): the sentinel must survive outside the active cells in both."""
    _, dms, mac = parity.make_bgc(33, 120, po, ragged=True, nColumns=111, with_dms=True, with_macros=True)
    parity.poison_outputs(dms); parity.poison_outputs(mac)
    a, b = dms.copy(), dms.copy()
    o.DMS_SourceSink(po, a); rt.DMS_SourceSink(rp, b)
    o.DMS_SurfaceFluxes(po, a); rt.DMS_SurfaceFluxes(rp, b)
    same(a.DMS_tendencies, b.DMS_tendencies, "DMS tendencies")
    for n in a.diag:
        same(a.diag[n], b.diag[n], "DMS " + n)
    for n in a.forcing:
        same(a.forcing[n], b.forcing[n], "DMS forcing " + n)
    for n in a.flux_diag:
        same(a.flux_diag[n], b.flux_diag[n], "DMS flux diag " + n)
    a, b = mac.copy(), mac.copy()
    o.MACROS_SourceSink(po, a); rt.MACROS_SourceSink(rp, b)
    same(a.MACROS_tendencies, b.MACROS_tendencies, "MACROS tendencies")
    for n in a.diag:
        same(a.diag[n], b.diag[n], "MACROS " + n)


def test_blocks_without_active_cells(po, rp):
    for case in ("numColumns=0", "kmax=0"):
        cols, dms, mac = parity.make_bgc(17, 50, po, with_dms=True, with_macros=True,
                                         nColumns=0 if case == "numColumns=0" else 50)
        if case == "kmax=0":
            for c in (cols, dms, mac):
                c.number_of_active_levels[:] = 0
        for c in (cols, dms, mac):
            parity.poison_outputs(c)
        cols.PH_PREV_3D[...] = 8.1
        a, b = cols.copy(), cols.copy()
        o.BGC_SourceSink(po, a, True); rt.BGC_SourceSink(rp, b, True)
        same_bgc(b, a, case)
        assert np.all(b.BGC_tendencies == 0.0) and np.all(b.PH_PREV_3D == 8.1)
        da, db = dms.copy(), dms.copy()
        o.DMS_SourceSink(po, da); rt.DMS_SourceSink(rp, db)
        same(da.DMS_tendencies, db.DMS_tendencies, case + " DMS")
        assert all(np.all(x == 7.25) for x in db.diag.values())


def test_the_comparison_has_teeth(po, rp):
    """The bit-exact comparison must notice a restatement that reads the source differently.  Two
    plausible misreadings, both produced with the oracle's own switches / inputs: (1) the
    single-precision literals of BGC_parms.F90:480-482 taken as doubles (what -fdefault-real-8
    would give, quirk Q6): epsTinv moves light_lim at the 1e-10 level; (2) T0_Kelvin_BGC left at 0
    instead of the host-set 273.15 (quirk Q7): it does not cancel in Tfunc in floating point."""
    cols, _, _ = parity.make_bgc(24, 64, po, ragged=True)
    ref = cols.copy()
    rt.BGC_SourceSink(rp, ref, True)
    a = cols.copy()
    o.BGC_SourceSink(po, a, True)
    assert np.array_equal(a.BGC_tendencies, ref.BGC_tendencies)
    p8 = o.Parms(default_real_8=True)
    b = cols.copy()
    o.BGC_SourceSink(p8, b, True)
    assert not np.array_equal(b.BGC_tendencies, ref.BGC_tendencies)
    d = np.abs(b.BGC_tendencies - ref.BGC_tendencies).max() / np.abs(ref.BGC_tendencies).max()
    assert 0 < d < 1e-6          # a last-digits effect: exactly what only a bit-exact check sees
    p0 = o.Parms()
    p0.bgc.T0_Kelvin_BGC = 0.0
    c = cols.copy()
    o.BGC_SourceSink(p0, c, True)
    assert not np.array_equal(c.BGC_tendencies, ref.BGC_tendencies)


# ------------------------------------------------------------------ properties of the reference itself
def test_no_result_depends_on_unset_memory(po):
    """libbgc_ref_poison.so fills every ALLOCATE with NaN bit patterns (gfortran leaves garbage):
    the outputs must not change, i.e. the reference reads no local array element it never set."""
    path = os.path.join(rt.REFDIR, "libbgc_ref_poison.so")
    if not os.path.exists(path):
        pytest.skip("poisoned build not present")
    cols, _, _ = parity.make_bgc(24, 96, po, ragged=True, nColumns=90)
    a, b = cols.copy(), cols.copy()
    rt.BGC_SourceSink(rt.RefParms(po), a, True)
    rt.BGC_SourceSink(rt.RefParms(po, L=rt.TLib(path, rt.META_PATH)), b, True)
    same_bgc(b, a, "poisoned allocations")


def test_columns_are_independent_in_the_reference(po, rp):
    """no routine reads another column.  Shown with the reference itself: a block
    computed in one call equals the same columns computed one call per column (numColumnsMax = 1,
    the way upstream MPAS drives the library), cold and warm - which is what makes contiguous
    column slabs per GPU, with no halo and no exchange, exact."""
    nL, nC = 30, 12
    cols, dms, mac = parity.make_bgc(nL, nC, po, ragged=True, with_dms=True, with_macros=True)
    whole, dwhole, mwhole = cols.copy(), dms.copy(), mac.copy()
    for _ in range(2):
        rt.BGC_SourceSink(rp, whole, True)
    rt.BGC_SurfaceFluxes(rp, whole)
    rt.DMS_SourceSink(rp, dwhole)
    rt.MACROS_SourceSink(rp, mwhole)
    for c in range(nC):
        one, d1, m1 = parity.make_bgc(nL, 1, po, ragged=True, column0=c, with_dms=True, with_macros=True)
        # same ragged depth as in the block (the generator keys everything on the global column)
        assert one.number_of_active_levels[0] == cols.number_of_active_levels[c]
        for _ in range(2):
            rt.BGC_SourceSink(rp, one, True)
        rt.BGC_SurfaceFluxes(rp, one)
        rt.DMS_SourceSink(rp, d1)
        rt.MACROS_SourceSink(rp, m1)
        same(one.BGC_tendencies[:, 0, :], whole.BGC_tendencies[:, c, :], "column %d tendencies" % c)
        same(one.PH_PREV_3D[:, 0], whole.PH_PREV_3D[:, c], "column %d pH" % c)
        same(one.forcing["netFlux"][0], whole.forcing["netFlux"][c], "column %d netFlux" % c)
        for n in ("diag_photoC", "diag_POC_FLUX_IN", "diag_CO3"):
            same(one.diag[n][:, 0], whole.diag[n][:, c], "column %d %s" % (c, n))
        same(d1.DMS_tendencies[:, 0, :], dwhole.DMS_tendencies[:, c, :], "column %d DMS" % c)
        same(m1.MACROS_tendencies[:, 0, :], mwhole.MACROS_tendencies[:, c, :], "column %d MACROS" % c)


def test_thread_local_module_state(po):
    """The reference is not re-entrant (solver state in module SAVE variables, co2calc.F90:65-67);
    the translation makes that state thread-local.  Slabs computed concurrently must equal the
    block computed in one call."""
    nL, nC = 30, 96
    cols, _, _ = parity.make_bgc(nL, nC, po, ragged=True)
    whole = cols.copy()
    rt.BGC_SourceSink(rt.RefParms(po), whole, True)
    slabs = []
    for c0 in range(0, nC, 24):
        s, _, _ = parity.make_bgc(nL, 24, po, ragged=True, column0=c0)
        slabs.append((s, None, None))
    run = rt.SlabRunner(po, 4)
    run.step(slabs)
    run.close()
    for i, (s, _, _) in enumerate(slabs):
        same(s.BGC_tendencies, whole.BGC_tendencies[:, 24 * i:24 * (i + 1), :], "slab %d" % i)


def test_default_parms_mean_the_defaults_after_a_perturbed_caller():
    """`f_qsw_par_DMS` gets its value in its declaration (DMS_parms.F90:191-192) and the `lrest_*`
    switches are module variables of BGC_mod (BGC_mod.F90:131-134): no `*_init` assigns them, so a
    caller that stored perturbed tables with sync_from (the differential fuzzer does, in the calling
    thread) used to leave them behind for every later RefParms of the process - the GPU shim test
    then compared the CUDA path with a reference running on another caller's PAR fraction.
    RefParms now puts the tunables back first (TLib.restore_tunables)."""
    res = {}

    def run():
        rt.RefParms(o.Parms())                      # this thread's first caller: the snapshot
        po2 = o.Parms()
        _retune(po2)
        po2.dms.f_qsw_par_DMS = 0.31
        po2.bgc.lrest_po4 = 1
        rt.RefParms(po2).sync_from(po2)
        assert rt.var("dms_parms__f_qsw_par_dms").value == 0.31
        po = o.Parms()
        rr = rt.RefParms(po)
        assert rt.var("dms_parms__f_qsw_par_dms").value == 0.45
        assert rt.var("bgc_mod__lrest_po4").value == 0
        cols, dms, mac = parity.make_bgc(20, 33, po, ragged=True, with_dms=True, with_macros=True)
        a, b, da, db = cols.copy(), cols.copy(), dms.copy(), dms.copy()
        o.BGC_SourceSink(po, a, True); rt.BGC_SourceSink(rr, b, True)
        o.DMS_SourceSink(po, da); rt.DMS_SourceSink(rr, db)
        res.update(a=a, b=b, da=da, db=db)

    import threading
    err = []

    def guarded():
        try:
            run()
        except BaseException as e:      # noqa: BLE001 - re-raised in the test's thread
            err.append(e)
    t = threading.Thread(target=guarded)
    t.start(); t.join()
    if err:
        raise err[0]
    assert np.array_equal(res["a"].BGC_tendencies, res["b"].BGC_tendencies, equal_nan=True)
    assert np.array_equal(res["da"].DMS_tendencies, res["db"].DMS_tendencies, equal_nan=True)
