"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded synthetic columns (tolerances: tests/parity.py), plus
size-independent properties at the full EC60to30 size.

Covers the edge cases the reference's own semantics create: ragged bathymetry, land
columns (kmax = 0), numColumns < numColumnsMax, odd numColumnsMax (cp.async
staging instead of the TMA bulk-copy staging), cold and warm pH brackets, diagnostics absent /
partially present, permuted tracer slots, dark columns, negative tracers, the
zero-mask of a functional group.
"""
import ctypes as C

import numpy as np
import pytest

import parity

pkg = parity.pkg
abi = pkg.abi
host = pkg.host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def o():
    return parity.oracle()


def _ctx(nL, nC, flavour=None, parms=None):
    parms = parms or host.Parms(flavour)
    return host.Context(nL, nC, device=0, flavour=flavour, parms=parms), parms


# ------------------------------------------------------------------ BGC_SourceSink
@pytest.mark.parametrize("nL,nC,nCols,ragged", [
    (60, 256, 256, False),    # two full blocks, bulk-copy path
    (60, 258, 250, True),     # partial last block, numColumns < numColumnsMax, land columns
    (33, 257, 257, True),     # odd numColumnsMax and level count -> cp.async staging instead of TMA bulk copies
    (60, 257, 255, True),     # odd numColumnsMax, even level count -> bulk copies of the aligned superset of a run
    (80, 130, 130, True),     # RRS18to6 level count
    (2, 64, 64, False),       # kmax forced to 1 below: level 1 is surface AND bottom
])
@pytest.mark.parametrize("device_mode", [True, False])
def test_bgc_source_sink_cold_and_warm(o, nL, nC, nCols, ragged, device_mode):
    ctx, parms = _ctx(nL, nC)
    po = o.Parms()
    cols, _, _ = parity.make_bgc(nL, nC, parms, nColumns=nCols, ragged=ragged)
    if nL == 2:
        cols.number_of_active_levels[:] = 1
    parity.poison_outputs(cols)
    ref = cols.copy()
    o.BGC_SourceSink(po, ref, True, nthreads=o.max_threads())
    got = parity.run_gpu_bgc(ctx, cols, device_mode=device_mode)
    parity.compare_bgc_source_sink(ref, got)
    # warm brackets: PH_PREV_* from the first pass
    ref2, got2 = ref.copy(), got.copy()
    o.BGC_SourceSink(po, ref2, True, nthreads=o.max_threads())
    got2 = parity.run_gpu_bgc(ctx, got2, device_mode=device_mode)
    parity.compare_bgc_source_sink(ref2, got2)
    st = ctx.status()
    assert st["no_bracket"] == 0 and st["no_convergence"] == 0 and st["nonfinite"] == 0, st
    ctx.close()


def test_bgc_strict_flavour_matches_too(o):
    ctx, parms = _ctx(40, 192, flavour="strict")
    po = o.Parms()
    cols, _, _ = parity.make_bgc(40, 192, parms, ragged=True)
    ref = cols.copy()
    o.BGC_SourceSink(po, ref, True, nthreads=o.max_threads())
    got = parity.run_gpu_bgc(ctx, cols, device_mode=True)
    errs = parity.compare_bgc_source_sink(ref, got, tol=1e-12)   # no FMA, IEEE divide: much tighter than 1e-10
    assert max(errs["tend[%d]" % (n + 1)] for n in range(30)) <= 1e-12
    ctx.close()


def test_alt_co2_use_eco_false_zeroes_dic_alt_tendency(o):
    ctx, parms = _ctx(30, 128)
    po = o.Parms()
    cols, _, _ = parity.make_bgc(30, 128, parms)
    ref = cols.copy()
    o.BGC_SourceSink(po, ref, False, nthreads=o.max_threads())
    got = parity.run_gpu_bgc(ctx, cols, device_mode=True, alt_co2_use_eco=False)
    parity.compare_bgc_source_sink(ref, got)
    assert np.all(got.BGC_tendencies[:, :, parms.ind.dic_alt_co2_ind - 1] == 0.0)
    ctx.close()


def test_diagnostics_absent_or_partial_do_not_change_tendencies():
    nL, nC = 40, 256
    ctx, parms = _ctx(nL, nC)
    cols, _, _ = parity.make_bgc(nL, nC, parms, ragged=True)
    full = parity.run_gpu_bgc(ctx, cols, device_mode=False)

    none = cols.copy()
    host.BGC_SourceSink(ctx, none, True, diagnostics=False)
    # three instantiations of the sweep (no / some / all diagnostics): the compiler contracts
    # FMAs differently in each, so equality holds to round-off, not bit for bit
    def same(a, b):
        return parity.nerr(a, b) <= 1e-13
    for n in range(30):
        assert same(none.BGC_tendencies[:, :, n], full.BGC_tendencies[:, :, n]), n
    assert np.array_equal(none.PH_PREV_3D, full.PH_PREV_3D)   # carbonate kernel: same code either way

    # a handful of diagnostics only (NULL-checked store path of the sweep)
    part = cols.copy()
    parity.poison_outputs(part)
    keep = ("diag_PAR_avg", "diag_photoC", "diag_Jint_Ctot", "diag_zsatcalc", "diag_CaCO3_form_zint", "diag_pH_3D")
    dg = abi.BgcDiagnostics()
    for n in keep:
        setattr(dg, n, abi.dptr(part.diag[n]))
    cin, cfo, cout = part.c_input(), part.c_forcing(), part.c_output()
    host.check(ctx.L, ctx.L.bgc_source_sink(ctx.ptr, C.byref(cin), C.byref(cfo), C.byref(cout), C.byref(dg),
                                            C.c_int(nL), C.c_int(nC), C.c_int(nC), C.c_int(1),
                                            C.c_int(abi.BGC_MEM_HOST_FORTRAN)))
    for n in range(30):
        assert same(part.BGC_tendencies[:, :, n], full.BGC_tendencies[:, :, n]), n
    for n in keep:
        if n == "diag_Jint_Ctot":
            continue    # ~0 residual of cancelling terms
        assert same(part.diag[n], full.diag[n]), n
    for n in part.diag:
        if n not in keep:
            assert np.all(part.diag[n] == 7.25), n   # untouched
    ctx.close()


def test_untouched_members_and_inactive_cells(o):
    """diag_POC_ACCUM / diag_DONr_remin / diag_DOPr_remin are never written (they are never
    written by the reference either); tendencies and every other BGC diagnostic are zero
    outside active cells; PH_PREV_* keep their values there (BGC_mod.F90:570, :625-727)."""
    nL, nC, nCols = 24, 128, 120
    ctx, parms = _ctx(nL, nC)
    cols, _, _ = parity.make_bgc(nL, nC, parms, nColumns=nCols, ragged=True)
    parity.poison_outputs(cols)
    cols.PH_PREV_3D[...] = 0.0
    cols.PH_PREV_ALT_CO2_3D[...] = 0.0
    mask = cols.active_mask()
    cols.PH_PREV_3D[~mask] = 3.5
    for dev in (True, False):
        got = parity.run_gpu_bgc(ctx, cols, device_mode=dev)
        for n in abi.BGC_DIAG_UNTOUCHED:
            assert np.all(got.diag[n] == 7.25), n
        assert np.all(got.BGC_tendencies[~mask] == 0.0)
        assert np.all(got.PH_PREV_3D[~mask] == 3.5)
        for n in abi.BGC_DIAG_K2:
            if n not in abi.BGC_DIAG_UNTOUCHED:
                assert np.all(got.diag[n][~mask] == 0.0), n
        for n in abi.BGC_DIAG_KA:
            assert np.all(got.diag[n][~mask] == 0.0), n
        land = np.nonzero(~mask.any(axis=0))[0]
        assert land.size > 0
        for n in abi.BGC_DIAG_C1:
            assert np.all(got.diag[n][land] == 0.0), n
    ctx.close()


def _retune(parms):
    """Non-default run-time parameters and a modified functional-group table: every branch that
    is keyed on a table value (not on the group's name) must follow."""
    b = parms.bgc
    b.parm_o2_min, b.parm_o2_min_delta = 6.0, 3.0
    b.parm_labile_ratio = 0.7
    b.parm_POMbury, b.parm_BSIbury = 1.3, 0.8
    b.parm_nitrif_par_lim = 2.5
    b.parm_kappa_nitrif *= 1.7
    b.parm_z_mort_0 *= 0.6
    b.parm_z_mort2_0 *= 1.4
    b.parm_fe_scavenge_rate0 *= 2.0
    b.parm_f_prod_sp_CaCO3 = 0.055
    b.parm_POC_diss, b.parm_SiO2_diss, b.parm_CaCO3_diss = 70.0e2, 300.0e2, 450.0e2
    for i in range(4):
        b.parm_scalelen_vals[i] *= (1.0 + 0.15 * i)
    b.lrest_no3 = b.lrest_po4 = b.lrest_sio3 = 1           # never set by the reference (Q8): the code path exists
    a = parms.autotrophs
    sp, diat, diaz, phaeo = a[parms.ind.sp_ind - 1], a[parms.ind.diat_ind - 1], a[parms.ind.diaz_ind - 1], a[parms.ind.phaeo_ind - 1]
    sp.temp_function, sp.temp_thresN, sp.temp_thresS, sp.temp_optN, sp.temp_optS = abi.DEFINES["BGC_TFNC_QUASI_MMRT"], 27.0, 26.0, 18.0, 17.0
    phaeo.temp_function, phaeo.temp_thres = abi.DEFINES["BGC_TFNC_Q10"], 1.0
    diat.Qp = 0.0061                                       # != Qp_zoo_pom: the remaining_P routing
    sp.kSiO3 = 0.4                                         # a silicate-limited group that carries no Si tracer
    diaz.graze_zoo, diaz.graze_poc, diaz.graze_doc = 0.25, 0.08, 0.10
    phaeo.grazee_ind = sp.grazee_ind                       # two groups share a grazer
    phaeo.Nfixer = 1                                       # a second N fixer: diag_Nfix of this group is no structural zero any more
    diat.agg_rate_max, diat.agg_rate_min, diat.mort2 = 0.7, 0.03, 0.012
    for g in (sp, diat, diaz, phaeo):
        g.PCref *= 1.1
        g.alphaPI *= 0.9


def test_non_default_parameters_and_group_table(o):
    nL, nC = 40, 512
    po = o.Parms()
    _retune(po)
    parms = host.Parms()
    _retune(parms)
    ctx, _ = _ctx(nL, nC, parms=parms)
    cols, _, _ = parity.make_bgc(nL, nC, parms, ragged=True, seed=0x5EED)
    rng = np.random.default_rng(11)
    cols.forcing["NUTR_RESTORE_RTAU"][...] = rng.uniform(0.0, 1.0e-6, size=(nL, nC))
    tr = cols.BGC_tracers
    cols.forcing["NO3_CLIM"][...] = tr[:, :, parms.ind.no3_ind - 1] * rng.uniform(0.8, 1.2, size=(nL, nC))
    cols.forcing["PO4_CLIM"][...] = tr[:, :, parms.ind.po4_ind - 1] * rng.uniform(0.8, 1.2, size=(nL, nC))
    cols.forcing["SiO3_CLIM"][...] = tr[:, :, parms.ind.sio3_ind - 1] * rng.uniform(0.8, 1.2, size=(nL, nC))
    parity.poison_outputs(cols)
    for device_mode in (True, False):
        ref = cols.copy()
        o.BGC_SourceSink(po, ref, True, nthreads=o.max_threads())
        got = parity.run_gpu_bgc(ctx, cols.copy(), device_mode=device_mode)
        parity.compare_bgc_source_sink(ref, got)
        assert np.abs(ref.diag["diag_NO3_RESTORE"]).max() > 0 and np.abs(ref.diag["diag_PO4_RESTORE"]).max() > 0
    st = ctx.status()
    assert st["no_bracket"] == 0 and st["no_convergence"] == 0 and st["nonfinite"] == 0, st
    ctx.close()


def test_permuted_tracer_slots(o):
    """The host chooses the tracer slots (BGC_indices_type): a permutation must only
    permute the outputs."""
    nL, nC = 30, 128
    rng = np.random.default_rng(7)
    perm = rng.permutation(30)
    ctx0, parms0 = _ctx(nL, nC)
    cols0, _, _ = parity.make_bgc(nL, nC, parms0, ragged=True)
    got0 = parity.run_gpu_bgc(ctx0, cols0, device_mode=True)
    ctx0.close()

    parms1 = host.Parms()
    parms1.permute_tracers(perm)
    ctx1 = host.Context(nL, nC, device=0, parms=parms1)
    cols1 = cols0.copy()
    for i in range(30):
        cols1.BGC_tracers[:, :, perm[i]] = cols0.BGC_tracers[:, :, i]
    got1 = parity.run_gpu_bgc(ctx1, cols1, device_mode=True)
    for i in range(30):
        assert np.array_equal(got1.BGC_tendencies[:, :, perm[i]], got0.BGC_tendencies[:, :, i]), i
    for n in ("diag_photoC", "diag_POC_REMIN", "diag_Jint_100m_Ntot"):
        assert np.array_equal(got1.diag[n], got0.diag[n]), n
    ctx1.close()


# ------------------------------------------------------------------ surface fluxes, co2calc
def test_surface_fluxes_and_side_effects(o):
    nL, nC = 20, 384
    ctx, parms = _ctx(nL, nC)
    po = o.Parms()
    parms.bgc.parm_Fe_bioavail = 0.5      # makes the in-place Fe scaling visible (BGC_mod.F90:2828-2838)
    po.bgc.parm_Fe_bioavail = 0.5
    ctx.set_params(parms)
    cols, _, _ = parity.make_bgc(nL, nC, parms, nColumns=380)
    cols.forcing["iceFraction"][:7] = [-0.2, 1.4, 0.3, 0.0, 1.0, 2.0, -1.0]
    for dev in (True, False):
        ref, got = cols.copy(), cols.copy()
        o.BGC_SurfaceFluxes(po, ref, nthreads=o.max_threads())
        if dev:
            d = host.DeviceBgcColumns(nL, nC, 380).load(got)
            host.BGC_SurfaceFluxes(ctx, d)
            ctx.synchronize()
            d.store(got)
        else:
            host.BGC_SurfaceFluxes(ctx, got)
        parity.compare_fields(ref.forcing, got.forcing, parity.TOL_TEND, "BGC_forcing",
                              solver_keys=("gasFlux", "netFlux", "surface_pH", "surface_pH_alt_co2"))
        parity.compare_fields(ref.flux_diag, got.flux_diag, parity.TOL_TEND, "BGC_flux_diagnostics",
                              solver_keys=parity.SOLVER_FLUX)
        assert np.array_equal(got.forcing["iceFraction"][:7], [0.0, 1.0, 0.3, 0.0, 1.0, 1.0, 0.0])
    ctx.close()


@pytest.mark.parametrize("o2,co2", [(0, 1), (1, 0), (0, 0)])
def test_surface_flux_switches(o, o2, co2):
    """lcalc_O2_gas_flux / lcalc_CO2_gas_flux off (BGC_mod.F90:2847, :2871) and
    lcalc_DMS_gas_flux off (DMS_mod.F90:846: the routine then does nothing at all)."""
    nL, nC = 12, 200
    ctx, parms = _ctx(nL, nC)
    po = o.Parms()
    cols, dms, _ = parity.make_bgc(nL, nC, parms, with_dms=True)
    cols.lcalc_O2_gas_flux, cols.lcalc_CO2_gas_flux = o2, co2
    dms.lcalc_DMS_gas_flux = 0
    parity.poison_outputs(cols)
    for a in dms.flux_diag.values():
        a[...] = 3.5
    dms.forcing["netFlux"][...] = 1.25
    for dev in (True, False):
        ref, got = cols.copy(), cols.copy()
        o.BGC_SurfaceFluxes(po, ref, nthreads=o.max_threads())
        dref, dgot = dms.copy(), dms.copy()
        o.DMS_SurfaceFluxes(po, dref)
        if dev:
            d = host.DeviceBgcColumns(nL, nC).load(got)
            dd = host.DeviceDmsColumns(nL, nC).load(dgot)
            host.BGC_SurfaceFluxes(ctx, d); host.DMS_SurfaceFluxes(ctx, dd)
            ctx.synchronize()
            d.store(got); dd.store(dgot)
        else:
            host.BGC_SurfaceFluxes(ctx, got); host.DMS_SurfaceFluxes(ctx, dgot)
        parity.compare_fields(ref.forcing, got.forcing, parity.TOL_TEND, "BGC_forcing",
                              solver_keys=("gasFlux", "netFlux", "surface_pH", "surface_pH_alt_co2"))
        parity.compare_fields(ref.flux_diag, got.flux_diag, parity.TOL_TEND, "BGC_flux_diagnostics",
                              solver_keys=parity.SOLVER_FLUX)
        assert np.array_equal(dgot.forcing["netFlux"], dref.forcing["netFlux"])
        for n in dref.flux_diag:
            assert np.array_equal(dgot.flux_diag[n], dref.flux_diag[n]), n
    ctx.close()


@pytest.mark.parametrize("warm", [False, True])
def test_co2calc_points(o, warm):
    n = 1 << 16
    ctx, _ = _ctx(1, 16)
    pts = pkg.synth_co2_points(n)
    if warm:
        cold = o.co2calc_points(pts, nthreads=o.max_threads())
        pts["phlo"] = cold["ph"] - 0.2
        pts["phhi"] = cold["ph"] + 0.2
    r = o.co2calc_points(pts, nthreads=o.max_threads())
    g = host.co2calc_points(ctx, pts)
    for k in ("ph", "co2star", "dco2star", "pco2surf", "dpco2"):
        assert parity.nerr(g[k], r[k]) <= parity.TOL_SOLVER, k
    # |dH| <= xacc (co2calc.F90:53)
    assert np.max(np.abs(10.0 ** -g["ph"] - 10.0 ** -r["ph"])) <= 1e-10
    assert host.co2calc_points(ctx, {k: v[:0] for k, v in pts.items()})["ph"].size == 0   # n = 0
    ctx.close()


# ------------------------------------------------------------------ DMS / MACROS
def test_co2calc_points_extreme_inputs(o):
    """Floors (dic_min, alk_min, salt_min: co2calc.F90:57-59, :843-846), brackets that do not
    contain the root (the growth loop, :920-938), polar and tropical temperatures, fresh water."""
    ctx, _ = _ctx(2, 64)
    n = 4096
    rng = np.random.default_rng(5)
    pts = pkg.synth_co2_points(n)
    pts["temp"] = rng.choice([-1.9, 0.0, 12.0, 30.0, 35.0], size=n)
    pts["salt"] = rng.choice([0.02, 0.5, 5.0, 20.0, 35.0, 41.0], size=n)     # 0.02 < salt_min
    pts["dic"] = rng.choice([0.5, 3.0, 800.0, 2000.0, 2600.0], size=n)       # 0.5, 3.0 < dic_min = 5.55
    pts["ta"] = pts["dic"] * rng.uniform(0.9, 1.4, size=n)
    pts["pt"] = rng.choice([0.0, 0.5, 5.0], size=n)
    pts["sit"] = rng.choice([0.0, 20.0, 200.0], size=n)
    lo = rng.choice([3.0, 6.0, 7.0, 9.5], size=n)                            # many brackets miss the root
    pts["phlo"], pts["phhi"] = lo, lo + rng.choice([0.4, 1.0, 2.0], size=n)
    r = o.co2calc_points(pts, nthreads=o.max_threads())
    g = host.co2calc_points(ctx, pts)
    ok = np.isfinite(r["ph"])
    assert ok.mean() > 0.9
    for k in ("ph", "co2star", "dco2star", "pco2surf", "dpco2"):
        assert parity.nerr(g[k][ok], r[k][ok]) <= parity.TOL_SOLVER, k
    assert np.array_equal(np.isfinite(g["ph"]), ok)
    ctx.close()


def test_block_without_active_cells(o):
    """numColumns = 0, and a block whose columns all have zero active levels: the whole-array zero
    fills of the reference (BGC_mod.F90:570, :625-727) are all that happens; PH_PREV_* stay."""
    nL, nC = 17, 130
    ctx, parms = _ctx(nL, nC)
    po = o.Parms()
    for case in ("numColumns=0", "kmax=0"):
        cols, dms, mac = parity.make_bgc(nL, nC, parms, with_dms=True, with_macros=True,
                                         nColumns=0 if case == "numColumns=0" else nC)
        if case == "kmax=0":
            cols.number_of_active_levels[:] = 0
            dms.number_of_active_levels[:] = 0
            mac.number_of_active_levels[:] = 0
        parity.poison_outputs(cols); parity.poison_outputs(dms); parity.poison_outputs(mac)
        cols.PH_PREV_3D[...] = 8.1
        for device_mode in (True, False):
            ref = cols.copy()
            o.BGC_SourceSink(po, ref, True)
            got = parity.run_gpu_bgc(ctx, cols.copy(), device_mode=device_mode)
            parity.compare_bgc_source_sink(ref, got)
            assert np.all(got.BGC_tendencies == 0.0) and np.all(got.PH_PREV_3D == 8.1), case
            assert np.all(got.diag["diag_POC_ACCUM"] == 7.25), case      # never touched by the reference
        dgot, mgot = dms.copy(), mac.copy()
        host.DMS_SourceSink(ctx, dgot); host.MACROS_SourceSink(ctx, mgot)
        assert np.all(dgot.DMS_tendencies == 0.0) and np.all(mgot.MACROS_tendencies == 0.0), case
        assert all(np.all(a == 7.25) for a in dgot.diag.values()), case   # DMS diagnostics are not zeroed
    ctx.close()


def _active(c):
    k = np.arange(1, c.nLevelsMax + 1)[:, None]
    kmax = c.number_of_active_levels.copy()
    kmax[c.nColumns:] = 0
    return k <= kmax[None, :]


@pytest.mark.parametrize("device_mode", [True, False])
def test_dms_and_macros(o, device_mode):
    nL, nC, nCols = 45, 258, 255
    ctx, parms = _ctx(nL, nC)
    po = o.Parms()
    _, dms, mac = parity.make_bgc(nL, nC, parms, nColumns=nCols, ragged=True, with_dms=True, with_macros=True)
    dref, mref = dms.copy(), mac.copy()
    o.DMS_SourceSink(po, dref, nthreads=o.max_threads())
    o.DMS_SurfaceFluxes(po, dref)
    o.MACROS_SourceSink(po, mref, nthreads=o.max_threads())
    dgot, mgot = dms.copy(), mac.copy()
    if device_mode:
        dd = host.DeviceDmsColumns(nL, nC, nCols).load(dgot)
        md = host.DeviceMacrosColumns(nL, nC, nCols).load(mgot)
        host.DMS_SourceSink(ctx, dd); host.DMS_SurfaceFluxes(ctx, dd); host.MACROS_SourceSink(ctx, md)
        ctx.synchronize()
        dd.store(dgot); md.store(mgot)
    else:
        host.DMS_SourceSink(ctx, dgot); host.DMS_SurfaceFluxes(ctx, dgot); host.MACROS_SourceSink(ctx, mgot)
    for n in range(abi.DMS_TRACER_CNT):
        assert parity.nerr(dgot.DMS_tendencies[:, :, n], dref.DMS_tendencies[:, :, n]) <= parity.TOL_TEND, n
    for n in range(abi.MACROS_TRACER_CNT):
        assert parity.nerr(mgot.MACROS_tendencies[:, :, n], mref.MACROS_tendencies[:, :, n]) <= parity.TOL_TEND, n
    # DMS / MACROS diagnostics are only defined on active cells (the reference never zeroes them)
    parity.compare_fields(dref.diag, dgot.diag, parity.TOL_TEND, "DMS diagnostics", mask=_active(dms))
    parity.compare_fields(mref.diag, mgot.diag, parity.TOL_TEND, "MACROS diagnostics", mask=_active(mac))
    cm = np.arange(nC) < nCols
    for nm in dref.flux_diag:
        assert parity.nerr(dgot.flux_diag[nm][cm], dref.flux_diag[nm][cm]) <= parity.TOL_TEND, nm
    assert parity.nerr(dgot.forcing["netFlux"][cm], dref.forcing["netFlux"][cm]) <= parity.TOL_TEND
    ctx.close()


@pytest.mark.parametrize("flavour", [None, "strict"])
def test_dms_kernel_choice_does_not_change_the_bits(o, flavour, monkeypatch):
    """DMS_SourceSink runs as a tile kernel (32 columns x all levels per block) or, for blocks of very many
    columns, as a column kernel (k_dms.cu: launch_dms_columns).  Which one a block gets depends on its
    width, so both must give the same tendencies and diagnostics bit for bit - a host that cuts its mesh
    differently must not see different numbers.  Both are also held against the oracle."""
    nL, nC, nCols = 45, 515, 511
    parms = host.Parms(flavour)
    _, dms, _ = parity.make_bgc(nL, nC, parms, nColumns=nCols, ragged=True, with_dms=True, with_macros=True)
    dref = dms.copy()
    o.DMS_SourceSink(o.Parms(), dref, nthreads=o.max_threads())
    outs = []
    for variant in ("1", "2"):   # 1: tile kernel, 2: column kernel (BGC_DMS_VARIANT is read when the ctx is created)
        monkeypatch.setenv("BGC_DMS_VARIANT", variant)
        ctx = host.Context(nL, nC, device=0, flavour=flavour, parms=parms)
        ctx.inventory_enable(True)      # each form writes its own number of block partials for the fold
        ctx.inventory_reset()
        got = dms.copy()
        dd = host.DeviceDmsColumns(nL, nC, nCols).load(got)
        host.DMS_SourceSink(ctx, dd)
        inv = ctx.inventory_get()
        ctx.synchronize()
        dd.store(got)
        ctx.close()
        dz = np.where(_active(dms), dms.cell_thickness, 0.0)
        want = np.einsum("kcn,kc->n", got.DMS_tendencies, dz)
        scale = np.einsum("kcn,kc->n", np.abs(got.DMS_tendencies), dz)
        assert np.all(np.abs(inv[30:30 + abi.DMS_TRACER_CNT] - want) <= 1e-12 * np.maximum(scale, 1e-300)), (variant, inv[30:44] - want)
        for n in range(abi.DMS_TRACER_CNT):
            assert parity.nerr(got.DMS_tendencies[:, :, n], dref.DMS_tendencies[:, :, n]) <= parity.TOL_TEND, (variant, n)
        parity.compare_fields(dref.diag, got.diag, parity.TOL_TEND, "DMS diagnostics, variant " + variant, mask=_active(dms))
        outs.append(got)
    a, b = outs
    assert np.array_equal(a.DMS_tendencies, b.DMS_tendencies)
    for nm in a.diag:
        assert np.array_equal(a.diag[nm], b.diag[nm]), nm


@pytest.mark.parametrize("ragged", [True, False])
def test_host_layout_pipeline_in_column_chunks(o, ragged, monkeypatch):
    """BGC_MEM_HOST_FORTRAN calls run as a two-slot pipeline over column chunks (bgc_capi.cu:
    host_pipeline).  Force tiny chunks so that several chunks, a partial last chunk, both
    slots and both diagnostic-upload modes (chunk fully active or not) are exercised."""
    monkeypatch.setenv("BGC_HOST_CHUNK_COLUMNS", "64")
    nL, nC = 40, 232
    nCols = 227 if ragged else nC
    ctx, parms = _ctx(nL, nC)
    po = o.Parms()
    cols, dms, mac = parity.make_bgc(nL, nC, parms, nColumns=nCols, ragged=ragged, with_dms=True, with_macros=True)
    parity.poison_outputs(cols)
    for d in (dms.diag, mac.diag):
        for a in d.values():
            a[...] = -777.0          # DMS / MACROS diagnostics keep the caller's values off active cells
    ref, dref, mref = cols.copy(), dms.copy(), mac.copy()
    o.BGC_SourceSink(po, ref, True, nthreads=o.max_threads())
    o.DMS_SourceSink(po, dref, nthreads=o.max_threads())
    o.MACROS_SourceSink(po, mref, nthreads=o.max_threads())
    got = parity.run_gpu_bgc(ctx, cols, device_mode=False)
    parity.compare_bgc_source_sink(ref, got)
    dgot, mgot = dms.copy(), mac.copy()
    host.DMS_SourceSink(ctx, dgot)
    host.MACROS_SourceSink(ctx, mgot)
    for n in range(abi.DMS_TRACER_CNT):
        assert parity.nerr(dgot.DMS_tendencies[:, :, n], dref.DMS_tendencies[:, :, n]) <= parity.TOL_TEND, n
    for n in range(abi.MACROS_TRACER_CNT):
        assert parity.nerr(mgot.MACROS_tendencies[:, :, n], mref.MACROS_tendencies[:, :, n]) <= parity.TOL_TEND, n
    act = _active(dms)
    parity.compare_fields(dref.diag, dgot.diag, parity.TOL_TEND, "DMS diagnostics", mask=act)
    parity.compare_fields(mref.diag, mgot.diag, parity.TOL_TEND, "MACROS diagnostics", mask=act)
    for nm, a in list(dgot.diag.items()) + list(mgot.diag.items()):
        assert np.all(a[~act] == -777.0), nm
    ctx.close()


def test_deferred_carbonate_join_gives_the_same_bits():
    """bgc_ctx_set_deferred_join moves the join of the carbonate side stream to the next join
    point; the results must be bit-identical to the strict (default) ordering."""
    import torch
    nL, nC = 60, 4096
    ctx, parms = _ctx(nL, nC)
    cols, dms, _ = parity.make_bgc(nL, nC, parms, ragged=True, with_dms=True)
    outs = []
    for deferred in (False, True):
        ctx.set_deferred_join(deferred)
        d = host.DeviceBgcColumns(nL, nC).load(cols)
        dd = host.DeviceDmsColumns(nL, nC).load(dms)
        for _ in range(2):                       # cold, then warm brackets
            host.BGC_SourceSink(ctx, d)
            host.BGC_SurfaceFluxes(ctx, d)
            host.DMS_SourceSink(ctx, dd)         # overlaps the carbonate kernel when deferred
        if deferred:
            ctx.carbonate_join()
        ctx.synchronize()
        outs.append({"tend": d.BGC_tendencies.clone(), "ph": d.PH_PREV_3D.clone(), "ph_alt": d.PH_PREV_ALT_CO2_3D.clone(),
                     "co3": d.diag["diag_CO3"].clone(), "zsat": d.diag["diag_zsatcalc"].clone(),
                     "zsata": d.diag["diag_zsatarag"].clone(), "dms": dd.DMS_tendencies.clone()})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k
    assert outs[0]["zsat"].abs().max().item() > 0
    # With the deferred join and no dms_source_sink behind it, the next join point is where the carbonate
    # kernel is waited for: same bits again.
    ctx.set_deferred_join(True)
    d = host.DeviceBgcColumns(nL, nC).load(cols)
    for _ in range(2):
        host.BGC_SourceSink(ctx, d)              # (the second call is itself a join point of the first)
        host.BGC_SurfaceFluxes(ctx, d)
    ctx.carbonate_join()
    ctx.synchronize()
    for k, v in (("tend", d.BGC_tendencies), ("ph", d.PH_PREV_3D), ("ph_alt", d.PH_PREV_ALT_CO2_3D),
                 ("co3", d.diag["diag_CO3"]), ("zsat", d.diag["diag_zsatcalc"]), ("zsata", d.diag["diag_zsatarag"])):
        assert torch.equal(outs[0][k], v), k
    ctx.close()


def test_device_side_diagnostics_accumulation(monkeypatch):
    """bgc_diag_accumulate_enable: host-layout calls keep returning tendencies every step but
    add the diagnostics into device accumulators; bgc_diag_flush returns scale * sum.  Must equal
    the average of the per-step diagnostics of the plain calls (several chunks, ragged columns)."""
    monkeypatch.setenv("BGC_HOST_CHUNK_COLUMNS", "64")
    nL, nC, nCols, steps = 30, 200, 190, 3
    ctx, parms = _ctx(nL, nC)
    cols, dms, mac = parity.make_bgc(nL, nC, parms, nColumns=nCols, ragged=True, with_dms=True, with_macros=True)
    act = _active(dms)
    # plain calls: per-step diagnostics (PH_PREV evolves cold -> warm, so the steps differ)
    a, ad, am = cols.copy(), dms.copy(), mac.copy()
    sums, tend_steps = {}, []
    for _ in range(steps):
        host.BGC_SourceSink(ctx, a); host.DMS_SourceSink(ctx, ad); host.MACROS_SourceSink(ctx, am)
        tend_steps.append(a.BGC_tendencies.copy())
        for pre, c_ in (("b.", a), ("d.", ad), ("m.", am)):
            for n, v in c_.diag.items():
                w = v if pre == "b." else np.where(act, v, 0.0)
                sums[pre + n] = sums.get(pre + n, 0.0) + w
    # accumulating calls
    b, bd, bm = cols.copy(), dms.copy(), mac.copy()
    for c_ in (b, bd, bm):
        for v in c_.diag.values():
            v[...] = -5.0
    ctx.diag_accumulate(True)
    for i in range(steps):
        host.BGC_SourceSink(ctx, b); host.DMS_SourceSink(ctx, bd); host.MACROS_SourceSink(ctx, bm)
        assert np.array_equal(b.BGC_tendencies, tend_steps[i])          # tendencies still come back every step
        assert np.all(b.diag["diag_PAR_avg"] == -5.0)                   # the caller's diagnostics are not touched
    ctx.diag_flush(b, bd, bm, scale=1.0 / steps, reset=True)
    ctx.diag_accumulate(False)
    never = ("diag_POC_ACCUM", "diag_DONr_remin", "diag_DOPr_remin")
    for pre, c_ in (("b.", b), ("d.", bd), ("m.", bm)):
        for n, v in c_.diag.items():
            if n in never:
                assert np.all(v == -5.0), n
                continue
            assert parity.nerr(v, sums[pre + n] / steps) <= 1e-14, (pre, n)
    ctx.close()


def test_mpas_tracer_layout_adapter():
    """bgc_layout_mpas_to_soa / bgc_layout_soa_to_mpas: T(tracer,k,cell) <-> SoA with a slot map,
    the reverse direction fused with the explicit tracer update (alpha = dt, beta = 1)."""
    import torch
    nT, nL, nC, nS = 33, 7, 1000 + 13, 30
    ctx, _ = _ctx(nL, nC)
    g = torch.Generator(device="cpu").manual_seed(7)
    mpas = torch.rand((nC, nL, nT), generator=g, dtype=torch.float64).cuda()      # (cell,k,n): n fastest
    perm = torch.randperm(nS, generator=g).tolist()
    slot = [0, 0, 0] + [p + 1 for p in perm]                                      # three tracers of another group
    soa = torch.full((nS, nL, nC), -1.0, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.mpas_to_soa(mpas.data_ptr(), soa.data_ptr(), slot, nL, nC)
    ctx.synchronize()
    for n in range(nT):
        if slot[n]:
            assert torch.equal(soa[slot[n] - 1], mpas[:, :, n].T.contiguous()), n
    # tendencies in SoA -> tracer update in MPAS layout
    tend = torch.rand((nS, nL, nC), generator=g, dtype=torch.float64).cuda()
    dt = 1800.0
    want = mpas.clone()
    for n in range(nT):
        if slot[n]:
            want[:, :, n] = want[:, :, n] + dt * tend[slot[n] - 1].T   # one multiply and one add: fma or not
    got = mpas.clone()
    torch.cuda.synchronize()
    ctx.soa_to_mpas(tend.data_ptr(), got.data_ptr(), slot, nL, nC, alpha=dt, beta=1.0)
    ctx.synchronize()
    assert torch.equal(got[:, :, :3], mpas[:, :, :3])                  # tracers outside the group untouched
    assert (got - want).abs().max().item() <= 1e-12 * want.abs().max().item()
    # beta = 0: pure layout change, bit exact round trip
    back = torch.zeros_like(mpas)
    ctx.soa_to_mpas(soa.data_ptr(), back.data_ptr(), slot, nL, nC, alpha=1.0, beta=0.0)
    ctx.synchronize()
    assert torch.equal(back[:, :, 3:], mpas[:, :, 3:])
    ctx.close()


def test_cuda_graph_replay_gives_the_same_bits():
    """bgc_graph_capture_begin/_end + bgc_graph_launch: one captured step (two streams, deferred
    carbonate join, inventory) replayed twice equals the same step issued call by call."""
    import torch
    nL, nC = 60, 4096
    ctx, parms = _ctx(nL, nC)
    ctx.inventory_enable(True)
    ctx.set_deferred_join(True)
    cols, dms, mac = parity.make_bgc(nL, nC, parms, ragged=True, with_dms=True, with_macros=True)
    d = host.DeviceBgcColumns(nL, nC).load(cols)
    dd = host.DeviceDmsColumns(nL, nC).load(dms)
    md = host.DeviceMacrosColumns(nL, nC).load(mac)

    def step():
        ctx.inventory_reset()
        host.BGC_SourceSink(ctx, d); host.BGC_SurfaceFluxes(ctx, d)
        host.DMS_SourceSink(ctx, dd); host.DMS_SurfaceFluxes(ctx, dd); host.MACROS_SourceSink(ctx, md)
        ctx.inventory_allreduce_begin()
    step(); step(); step()                      # cold + two warm steps, call by call
    inv_ref = ctx.inventory_allreduce_end()
    ref = {"tend": d.BGC_tendencies.clone(), "ph": d.PH_PREV_3D.clone(), "co3": d.diag["diag_CO3"].clone(),
           "dms": dd.DMS_tendencies.clone(), "mac": md.MACROS_tendencies.clone(), "net": d.forcing["netFlux"].clone()}
    n_calls = ctx.launch_count()
    # same state again, this time: cold + warm call by call, then capture and replay
    d.load(cols); dd.load(dms); md.load(mac)
    step()
    ctx.synchronize()
    ctx.timing_reset()
    ctx.graph_capture_begin()
    step()
    g = ctx.graph_capture_end()
    assert ctx.launch_count() == 0              # captured, not executed
    ctx.graph_launch(g); ctx.graph_launch(g)
    inv = ctx.inventory_allreduce_end()
    assert ctx.launch_count() > 0
    got = {"tend": d.BGC_tendencies, "ph": d.PH_PREV_3D, "co3": d.diag["diag_CO3"], "dms": dd.DMS_tendencies,
           "mac": md.MACROS_tendencies, "net": d.forcing["netFlux"]}
    for k in ref:
        assert torch.equal(ref[k], got[k]), k
    assert np.array_equal(inv, inv_ref)
    ctx.graph_destroy(g)
    ctx.close()
    assert n_calls > 0


def test_zero_biomass_shortcut_changes_nothing():
    """bgc_ctx_set_zero_shortcut: skipping the body of a functional group whose biomass is zero in
    a whole warp must give the same tendencies bit for bit, and the same diagnostics (the four
    nutrient-limitation diagnostics are recomputed in the shortcut: last-bit FMA differences allowed)."""
    import torch
    nL, nC = 60, 2048 + 64
    ctx, parms = _ctx(nL, nC)
    ctx.inventory_enable(True)
    cols, _, _ = parity.make_bgc(nL, nC, parms, ragged=True)
    outs = []
    for on in (True, False):
        ctx.set_zero_shortcut(on)
        d = host.DeviceBgcColumns(nL, nC).load(cols)
        ctx.inventory_reset()
        host.BGC_SourceSink(ctx, d)
        host.BGC_SourceSink(ctx, d)
        inv = ctx.inventory_get()
        ctx.synchronize()
        outs.append((d, inv))
    (a, inv_a), (b, inv_b) = outs
    assert torch.equal(a.BGC_tendencies, b.BGC_tendencies)
    assert torch.equal(a.PH_PREV_3D, b.PH_PREV_3D)
    assert np.array_equal(inv_a, inv_b)
    masked_cells = (a.BGC_tracers[16] == 0).float().mean().item()      # spC == 0: the shortcut had work to skip
    assert masked_cells > 0.3, masked_cells
    for n, v in a.diag.items():
        w = b.diag[n]
        if n in ("diag_N_lim", "diag_P_lim", "diag_Fe_lim", "diag_SiO3_lim"):
            assert (v - w).abs().max().item() <= 4e-16, n
        else:
            assert torch.equal(v, w), n
    ctx.close()


def test_zero_biomass_shortcut_lets_nonfinite_inputs_through():
    """0 * NaN is NaN: where the reference multiplies a zeroed functional group with a non-finite factor
    (a NaN temperature, an infinite nutrient) its tendencies are NaN, and the model sees them.  The shortcut
    writes the zeros of a group body only where every factor of the level is finite in the whole warp
    (k_eco.cu: lvl_finite): with such inputs in deep cells whose biomass is exactly zero, shortcut on and
    off give the same bits, NaN for NaN."""
    nL, nC = 60, 512
    ctx, parms = _ctx(nL, nC)
    cols, _, _ = parity.make_bgc(nL, nC, parms)
    k = nL - 5
    assert (cols.BGC_tracers[k, :, parms.ind.spC_ind - 1] == 0).all()   # deep ocean of the synthetic columns: no biomass
    assert cols.number_of_active_levels[[40, 300]].min() > k
    bad = cols.copy()
    bad.PotentialTemperature[k, 40] = np.nan
    bad.BGC_tracers[k, 300, parms.ind.no3_ind - 1] = np.inf
    outs = []
    for on in (True, False):
        ctx.set_zero_shortcut(on)
        outs.append(parity.run_gpu_bgc(ctx, bad.copy(), device_mode=True))
        ctx.status(reset=True)
    a, b = outs
    assert np.array_equal(np.isnan(a.BGC_tendencies), np.isnan(b.BGC_tendencies))
    assert np.array_equal(np.nan_to_num(a.BGC_tendencies, nan=1.5, posinf=2.5, neginf=-2.5),
                          np.nan_to_num(b.BGC_tendencies, nan=1.5, posinf=2.5, neginf=-2.5))
    for n in a.diag:
        assert np.array_equal(np.isnan(a.diag[n]), np.isnan(b.diag[n])), n
    assert np.isnan(a.BGC_tendencies[k, 40, :]).any()          # the NaN temperature reaches the zeroed groups
    assert not np.isfinite(a.BGC_tendencies[k, 300, :]).all()   # so does the infinite nitrate
    ctx.close()


# ------------------------------------------------------------------ inventory
def test_inventory_vector_matches_the_outputs():
    nL, nC, nCols = 36, 514, 500
    ctx, parms = _ctx(nL, nC)
    cols, dms, mac = parity.make_bgc(nL, nC, parms, nColumns=nCols, ragged=True, with_dms=True, with_macros=True)
    ctx.inventory_enable(True)
    ctx.inventory_reset()
    host.BGC_SourceSink(ctx, cols); host.DMS_SourceSink(ctx, dms); host.MACROS_SourceSink(ctx, mac)
    inv = ctx.inventory_allreduce()     # single rank: a copy
    mask = cols.active_mask()
    dz = np.where(mask, cols.cell_thickness, 0.0)
    want = np.zeros(abi.BGC_INVENTORY_LEN)
    want[0:30] = np.einsum("kcn,kc->n", cols.BGC_tendencies, dz)
    want[30:44] = np.einsum("kcn,kc->n", dms.DMS_tendencies, dz)
    want[44:52] = np.einsum("kcn,kc->n", mac.MACROS_tendencies, dz)
    for i, nm in enumerate(("Ctot", "100m_Ctot", "Ntot", "100m_Ntot", "Ptot", "100m_Ptot", "Sitot", "100m_Sitot")):
        want[52 + i] = cols.diag["diag_Jint_" + nm][:nCols].sum()
    want[60], want[61] = mask.sum(), mask.any(axis=0).sum()
    scale = np.zeros_like(want)
    scale[0:30] = np.einsum("kcn,kc->n", np.abs(cols.BGC_tendencies), dz)
    scale[30:44] = np.einsum("kcn,kc->n", np.abs(dms.DMS_tendencies), dz)
    scale[44:52] = np.einsum("kcn,kc->n", np.abs(mac.MACROS_tendencies), dz)
    scale[52:60] = np.abs(cols.diag["diag_Jint_100m_Ctot"][:nCols]).sum() + 1.0
    scale[60:62] = 1.0
    assert np.all(np.abs(inv - want)[:62] <= 1e-12 * np.maximum(scale[:62], 1e-300)), (inv - want)[:62]
    # accumulates across calls until reset; bit-reproducible (no atomics)
    host.BGC_SourceSink(ctx, cols)
    inv2 = ctx.inventory_get()
    ctx.inventory_reset()
    host.BGC_SourceSink(ctx, cols)
    inv3 = ctx.inventory_get()
    assert np.allclose(inv2[:30], inv[:30] + inv3[:30], rtol=1e-13, atol=0)
    ctx.inventory_reset()
    host.BGC_SourceSink(ctx, cols)
    assert np.array_equal(ctx.inventory_get(), inv3)
    # single rank: the all-reduce is the identity; the stream-ordered form returns the same vector
    assert np.array_equal(ctx.inventory_allreduce(), inv3)
    ctx.inventory_allreduce_begin()
    assert np.array_equal(ctx.inventory_allreduce_end(), inv3)
    ctx.close()


# ------------------------------------------------------------------ full size: properties
def test_ec60to30_full_size_properties():
    """235 160 columns x 60 levels, device resident.  The oracle would need minutes here,
    so check what the algorithm guarantees: element conservation (the code's own Jint
    integrals vanish to round-off relative to their upper-100 m parts, BGC_mod.F90:1875-1938),
    slab independence (a sub-slab run alone reproduces the same bits - the sharding
    property), clean solver status, warm-bracket idempotence of pH."""
    import torch
    nL, nC = 60, 235160
    ctx, parms = _ctx(nL, nC)
    d = host.DeviceBgcColumns(nL, nC)
    # fill 8 slabs through a small host container (keeps host memory modest)
    slab = 29395
    assert slab * 8 == nC
    h = pkg.BgcColumns(nL, slab)
    for s in range(8):
        pkg.synth_fill(h, bgc_ind=parms.ind, column0=s * slab, ragged=True)
        sl = slice(s * slab, (s + 1) * slab)
        d.BGC_tracers[:, :, sl] = torch.from_numpy(np.ascontiguousarray(np.transpose(h.BGC_tracers, (2, 0, 1))))
        for n in d.K2_IN:
            getattr(d, n)[:, sl] = torch.from_numpy(np.ascontiguousarray(getattr(h, n)))
        d.cell_latitude[sl] = torch.from_numpy(h.cell_latitude)
        d.number_of_active_levels[sl] = torch.from_numpy(h.number_of_active_levels)
        d.forcing["FESEDFLUX"][:, sl] = torch.from_numpy(np.ascontiguousarray(h.forcing["FESEDFLUX"]))
        for n in ("dust_FLUX_IN", "ShortWaveFlux_surface"):
            d.forcing[n][sl] = torch.from_numpy(h.forcing[n])
    torch.cuda.synchronize()
    host.BGC_SourceSink(ctx, d)           # cold
    ctx.synchronize()
    ph_cold = d.PH_PREV_3D.clone()
    host.BGC_SourceSink(ctx, d)           # warm
    ctx.synchronize()
    st = ctx.status()
    assert st["no_bracket"] == 0 and st["no_convergence"] == 0, st
    kmax = d.number_of_active_levels.long()
    active = torch.arange(nL, device=d.device)[:, None] < kmax[None, :]
    assert torch.isfinite(d.BGC_tendencies).all()
    assert (d.BGC_tendencies[:, ~active] == 0).all()
    # the warm-bracket solve lands on the same root: both stop within xacc = 1e-10 mol/kg of it
    # (co2calc.F90:53), whatever that means in pH units at that H+
    dH = (10.0 ** -d.PH_PREV_3D - 10.0 ** -ph_cold)[active].abs().max().item()
    assert dH <= 2.0e-10, dH
    assert (d.PH_PREV_3D - ph_cold)[active].abs().median().item() < 1e-3
    for el in ("C", "N", "P", "Si"):
        tot = d.diag["diag_Jint_%stot" % el].abs().max().item()
        part = d.diag["diag_Jint_100m_%stot" % el].abs().max().item()
        assert tot <= 1e-10 * part, (el, tot, part)
    # slab independence: columns [2*slab, 3*slab) alone, bit for bit.  The slab is 29 395 columns
    # wide - ODD, so its level rows are not 16-byte aligned and the sweep stages them with 8-byte
    # cp.async copies instead of TMA bulk copies: the staging differs, the arithmetic does not.
    sl = slice(2 * slab, 3 * slab)
    sub = host.DeviceBgcColumns(nL, slab)
    own = slice(0, slab)
    sub.BGC_tracers[:, :, own].copy_(d.BGC_tracers[:, :, sl])
    for n in d.K2_IN:
        getattr(sub, n)[:, own].copy_(getattr(d, n)[:, sl])
    sub.cell_latitude[own].copy_(d.cell_latitude[sl])
    sub.number_of_active_levels[own].copy_(d.number_of_active_levels[sl])
    sub.forcing["FESEDFLUX"][:, own].copy_(d.forcing["FESEDFLUX"][:, sl])
    for n in ("dust_FLUX_IN", "ShortWaveFlux_surface"):
        sub.forcing[n][own].copy_(d.forcing[n][sl])
    sub.PH_PREV_3D[:, own].copy_(ph_cold[:, sl])
    sub.PH_PREV_ALT_CO2_3D[:, own].copy_(ph_cold[:, sl])   # both solves share inputs (BGC_mod.F90:975): same root
    ctx2 = host.Context(nL, slab, device=0, parms=parms)
    torch.cuda.synchronize()   # the copies above ran on torch's stream, the ctx has its own
    host.BGC_SourceSink(ctx2, sub)
    ctx2.synchronize()
    assert torch.equal(sub.BGC_tendencies[:, :, own], d.BGC_tendencies[:, :, sl])
    for nm in ("diag_POC_REMIN", "diag_photoC", "diag_Jint_Ctot", "diag_CO3", "diag_zsatcalc"):
        assert torch.equal(sub.diag[nm][..., own], d.diag[nm][..., sl]), nm
    assert torch.equal(sub.PH_PREV_3D[:, own], d.PH_PREV_3D[:, sl])
    ctx2.close()
    ctx.close()


# ------------------------------------------------------------------ error behaviour
def test_status_counts_nonfinite_cells():
    """BgcStatus.nonfinite: a NaN that reaches a tendency is counted (the reference has no error
    reporting at all); the other columns are unaffected."""
    nL, nC = 20, 256
    ctx, parms = _ctx(nL, nC)
    cols, _, _ = parity.make_bgc(nL, nC, parms)
    clean = parity.run_gpu_bgc(ctx, cols.copy(), device_mode=True)
    assert ctx.status(reset=True)["nonfinite"] == 0
    bad = cols.copy()
    bad.PotentialTemperature[3, 17] = np.nan
    bad.BGC_tracers[5, 101, parms.ind.doc_ind - 1] = np.inf
    got = parity.run_gpu_bgc(ctx, bad, device_mode=True)
    st = ctx.status(reset=True)
    assert st["nonfinite"] >= 2, st
    keep = np.ones(nC, bool); keep[[17, 101]] = False
    assert np.array_equal(got.BGC_tendencies[:, keep, :], clean.BGC_tendencies[:, keep, :])
    assert ctx.status()["nonfinite"] == 0
    ctx.close()


def test_no_kernel_writes_outside_its_arrays(monkeypatch):
    """Guard bands: every device array of the containers is carved out of one pool with 64
    sentinel doubles before and after it; no kernel may touch a sentinel (partial last block,
    ragged columns, numColumns < numColumnsMax, odd and even extents, inventory on).
    (compute-sanitizer is not available on the GPU pool: this is the in-tree substitute for
    out-of-bounds stores.)"""
    import torch
    GUARD, SENT = 64, -123456.5
    pool = torch.full((40_000_000,), SENT, dtype=torch.float64, device="cuda")
    ipool = torch.full((200_000,), -77, dtype=torch.int32, device="cuda")
    state = {"off": GUARD, "ioff": GUARD, "spans": [], "ispans": []}

    def alloc(self, shape, dtype=None):
        n = int(np.prod(shape))
        if dtype is not None and dtype != torch.float64:
            a = ipool[state["ioff"]:state["ioff"] + n]
            state["ispans"].append((state["ioff"], n)); state["ioff"] += n + GUARD
            a.zero_()
            return a.view(shape)
        off = (state["off"] + 1) & ~1          # keep 16-byte alignment (bulk-copy path)
        a = pool[off:off + n]
        state["spans"].append((off, n)); state["off"] = off + n + GUARD
        a.zero_()
        return a.view(shape)
    monkeypatch.setattr(host._DeviceMixin, "_alloc", alloc)
    for nL, nC, nCols in ((9, 300, 290), (7, 129, 129)):
        ctx, parms = _ctx(nL, nC)
        ctx.inventory_enable(True)
        cols, dms, mac = parity.make_bgc(nL, nC, parms, nColumns=nCols, ragged=True, with_dms=True, with_macros=True)
        d = host.DeviceBgcColumns(nL, nC, nCols).load(cols)
        dd = host.DeviceDmsColumns(nL, nC, nCols).load(dms)
        md = host.DeviceMacrosColumns(nL, nC, nCols).load(mac)
        for _ in range(2):
            host.BGC_SourceSink(ctx, d); host.BGC_SurfaceFluxes(ctx, d)
            host.DMS_SourceSink(ctx, dd); host.DMS_SurfaceFluxes(ctx, dd); host.MACROS_SourceSink(ctx, md)
        ctx.inventory_allreduce()
        ctx.synchronize()
        ctx.close()
    mask = torch.ones_like(pool, dtype=torch.bool)
    for off, n in state["spans"]:
        mask[off:off + n] = False
    assert bool((pool[mask] == SENT).all()), "a kernel wrote outside its FP64 arrays"
    imask = torch.ones_like(ipool, dtype=torch.bool)
    for off, n in state["ispans"]:
        imask[off:off + n] = False
    assert bool((ipool[imask] == -77).all())
    assert len(state["spans"]) > 300


def test_error_codes():
    L = host.lib()
    ctx = host.Context(8, 32, device=0)          # no parameters yet
    cols = pkg.BgcColumns(8, 32)
    with pytest.raises(host.BgcError, match="set_params"):
        host.BGC_SourceSink(ctx, cols)
    parms = host.Parms()
    ctx.set_params(parms)
    cin, cfo, cout, cdg = cols.c_input(), cols.c_forcing(), cols.c_output(), cols.c_diag()
    cin.cell_thickness = abi.dptr(None)
    rc = L.bgc_source_sink(ctx.ptr, C.byref(cin), C.byref(cfo), C.byref(cout), C.byref(cdg), C.c_int(8), C.c_int(32),
                           C.c_int(32), C.c_int(1), C.c_int(abi.BGC_MEM_HOST_FORTRAN))
    assert rc == abi.DEFINES["BGC_ERR_ARG"]
    rc = L.bgc_source_sink(ctx.ptr, C.byref(cols.c_input()), C.byref(cfo), C.byref(cout), C.byref(cdg), C.c_int(8),
                           C.c_int(32), C.c_int(33), C.c_int(1), C.c_int(abi.BGC_MEM_HOST_FORTRAN))
    assert rc == abi.DEFINES["BGC_ERR_ARG"]       # numColumns > numColumnsMax
    ctx.close()


# ------------------------------------------------------------------ round 2
@pytest.mark.parametrize("nL", [37, 36])
@pytest.mark.parametrize("nC_a,nC_b", [(257, 258), (301, 512), (1, 2), (511, 513)])
def test_block_width_does_not_change_the_bits(nC_a, nC_b, nL):
    """The same columns computed in blocks of different numColumnsMax - even (TMA bulk copies of the runs), odd
    with an even level count (bulk copies of the aligned superset of every other level's run, the last block's
    last level rounded down with its last column fetched by its own thread), odd with an odd level count (8-byte
    cp.async staging), one block and several - give identical bits in every output: there is ONE instantiation
    of the sweep's arithmetic (k_eco.cu)."""
    import torch
    parms = host.Parms()
    n = min(nC_a, nC_b)
    base, _, _ = parity.make_bgc(nL, n, parms, ragged=True)
    outs = []
    for nC in (nC_a, nC_b):
        cols = pkg.BgcColumns(nL, nC, n)
        for nm in ("BGC_tracers", "PotentialTemperature", "Salinity", "cell_center_depth", "cell_thickness",
                   "cell_bottom_depth"):
            getattr(cols, nm)[:, :n] = getattr(base, nm)
        cols.cell_latitude[:n] = base.cell_latitude
        cols.number_of_active_levels[:n] = base.number_of_active_levels
        for nm, a in base.forcing.items():
            if a.ndim == 2 and a.shape[0] == nL:
                cols.forcing[nm][:, :n] = a
            else:
                cols.forcing[nm][:n] = a
        ctx = host.Context(nL, nC, device=0, parms=parms)
        got = parity.run_gpu_bgc(ctx, cols, device_mode=True)
        got = parity.run_gpu_bgc(ctx, got, device_mode=True)      # warm pass as well
        outs.append(got)
        ctx.close()
    a, b = outs
    assert np.array_equal(a.BGC_tendencies[:, :n], b.BGC_tendencies[:, :n])
    assert np.array_equal(a.PH_PREV_3D[:, :n], b.PH_PREV_3D[:, :n])
    for nm in a.diag:
        x, y = a.diag[nm], b.diag[nm]
        if x.ndim >= 2 and x.shape[0] == nL:
            assert np.array_equal(x[:, :n], y[:, :n]), nm
        else:
            assert np.array_equal(x[:n], y[:n]), nm


@pytest.mark.parametrize("share", [0, 1, 100])
def test_carbonate_placement_does_not_change_the_bits(share, monkeypatch):
    """bgc_source_sink places the carbonate kernel by the size of the sweep (bgc_capi.cu: same stream,
    beside the sweep, behind it, or - for a sweep of less than one wave - confined to the SMs the sweep
    leaves idle, as SM-filling persistent blocks over a share of the cells with the rest behind the
    sweep).  Every placement gives the same bits: share = 1 % (of the modelled capacity of the idle SMs) cuts this mesh in the middle of a level,
    100 % puts all of it into the persistent blocks."""
    nL, nC = 60, 2300
    parms = host.Parms()
    cols, _, _ = parity.make_bgc(nL, nC, parms, ragged=True)
    if share:
        monkeypatch.setenv("BGC_CO3_SHARE", str(share))
    outs = []
    for mode in ((0, 3) if share else (0, 1, 2, 3)):
        ctx = host.Context(nL, nC, device=0, parms=parms)
        ctx.set_concurrency(mode)
        got = parity.run_gpu_bgc(ctx, cols, device_mode=True)
        got = parity.run_gpu_bgc(ctx, got, device_mode=True)      # warm pass as well
        outs.append(got)
        ctx.close()
    a = outs[0]
    for b in outs[1:]:
        assert np.array_equal(a.BGC_tendencies, b.BGC_tendencies)
        assert np.array_equal(a.PH_PREV_3D, b.PH_PREV_3D)
        assert np.array_equal(a.PH_PREV_ALT_CO2_3D, b.PH_PREV_ALT_CO2_3D)
        for nm in a.diag:
            assert np.array_equal(a.diag[nm], b.diag[nm]), nm


def test_comp_co3terms_and_sat_vals_points(o):
    """The rest of the co2calc module's public trio on the GPU (co2calc.F90:214-316, :1096-1238), batched,
    against the oracle point by point: surface and deep levels, cold and warm brackets."""
    ctx, parms = _ctx(2, 64)
    n = 6000
    pts = pkg.synth_co2_points(n)
    rng = np.random.default_rng(11)
    k = rng.integers(1, 81, n).astype(np.int32)
    k[:100] = 1
    depth = np.where(k == 1, 5.0, rng.uniform(10.0, 6000.0, n))
    for warm in (False, True):
        lo, hi = (pts["phlo"] - 1.0, pts["phhi"]) if not warm else (pts["phlo"], pts["phhi"])
        ref = {nm: np.zeros(n) for nm in ("pH", "H2CO3", "HCO3", "CO3")}
        sat_c, sat_a = np.zeros(n), np.zeros(n)
        for i in range(n):
            r = o.comp_CO3terms(int(k[i]), depth[i], pts["temp"][i], pts["salt"][i], pts["dic"][i], pts["ta"][i],
                                pts["pt"][i], pts["sit"][i], lo[i], hi[i])
            for nm in ref:
                ref[nm][i] = r[nm]
            sat_c[i], sat_a[i] = o.co3_sat_vals(int(k[i]), depth[i], pts["temp"][i], pts["salt"][i])
        got = host.comp_CO3terms_points(ctx, k, depth, pts["temp"], pts["salt"], pts["dic"], pts["ta"], pts["pt"],
                                        pts["sit"], lo, hi)
        for nm in ref:
            assert parity.nerr(got[nm], ref[nm]) <= parity.TOL_SOLVER, (warm, nm)
        assert np.max(np.abs(10.0 ** -got["pH"] - 10.0 ** -ref["pH"])) <= 1e-10
        gc, ga = host.comp_co3_sat_vals_points(ctx, k, depth, pts["temp"], pts["salt"])
        assert parity.nerr(gc, sat_c) <= parity.TOL_TEND and parity.nerr(ga, sat_a) <= parity.TOL_TEND
        pts["phlo"], pts["phhi"] = ref["pH"] - 0.2, ref["pH"] + 0.2
    # scalar level index for the whole batch
    g1 = host.comp_co3_sat_vals_points(ctx, 1, depth, pts["temp"], pts["salt"])
    g2 = host.comp_co3_sat_vals_points(ctx, np.ones(n, np.int32), depth, pts["temp"], pts["salt"])
    assert np.array_equal(g1[0], g2[0]) and np.array_equal(g1[1], g2[1])
    st = ctx.status()
    assert st["no_bracket"] == 0 and st["no_convergence"] == 0, st
    ctx.close()


class _RawDevice:
    """a raw device address as something torch.as_tensor understands"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


def test_resident_state_and_restart_round_trip(o):
    """bgc_state_*: PH_PREV_3D / PH_PREV_ALT_CO2_3D / surface_pH / surface_pH_alt_co2 (the fields the
    reference carries between steps and writes to restart files, BGC_parms.F90:151-152, :170-171) live
    in the ctx; a device-resident model passes the state pointers every step.  A run that is stopped
    after two steps, written out with bgc_state_get and restarted in a NEW ctx with bgc_state_set
    continues with the bits of the uninterrupted run; the reference agrees at the usual tolerances."""
    import torch
    nL, nC = 31, 200
    parms = host.Parms()
    po = o.Parms()
    cols, _, _ = parity.make_bgc(nL, nC, parms, ragged=True)

    def resident(ctx):
        dev = host.DeviceBgcColumns(nL, nC).load(cols)
        dev.PH_PREV_3D = torch.as_tensor(_RawDevice(ctx.state_device_ptr("PH_PREV_3D", nL, nC), (nL, nC)), device="cuda")
        dev.PH_PREV_ALT_CO2_3D = torch.as_tensor(_RawDevice(ctx.state_device_ptr("PH_PREV_ALT_CO2_3D", nL, nC), (nL, nC)),
                                                 device="cuda")
        dev.forcing["surface_pH"] = torch.as_tensor(_RawDevice(ctx.state_device_ptr("surface_pH", nL, nC), (nC,)), device="cuda")
        dev.forcing["surface_pH_alt_co2"] = torch.as_tensor(
            _RawDevice(ctx.state_device_ptr("surface_pH_alt_co2", nL, nC), (nC,)), device="cuda")
        return dev

    def step(ctx, dev):
        host.BGC_SourceSink(ctx, dev, True, True)
        host.BGC_SurfaceFluxes(ctx, dev)
        ctx.synchronize()

    ctx = host.Context(nL, nC, device=0, parms=parms)
    dev = resident(ctx)
    ctx.synchronize()
    assert float(dev.PH_PREV_3D.abs().max()) == 0.0          # created zero-filled: "no previous pH"
    step(ctx, dev); step(ctx, dev)
    names = ("PH_PREV_3D", "PH_PREV_ALT_CO2_3D", "surface_pH", "surface_pH_alt_co2")
    saved = {nm: ctx.state_get(nm, nL, nC) for nm in names}          # what a restart file holds
    assert np.array_equal(saved["PH_PREV_3D"], dev.PH_PREV_3D.cpu().numpy())
    assert np.array_equal(saved["surface_pH"], dev.forcing["surface_pH"].cpu().numpy())
    step(ctx, dev)                                                   # step 3 of the uninterrupted run
    want = dev.store(cols.copy())
    ctx.close()

    ctx2 = host.Context(nL, nC, device=0, parms=parms)                 # the restarted run
    dev2 = resident(ctx2)
    for nm in names:
        ctx2.state_set(nm, saved[nm], nL, nC)
        assert np.array_equal(ctx2.state_get(nm, nL, nC), saved[nm])
    step(ctx2, dev2)
    got = dev2.store(cols.copy())
    assert np.array_equal(got.BGC_tendencies, want.BGC_tendencies)
    assert np.array_equal(got.PH_PREV_3D, want.PH_PREV_3D)
    assert np.array_equal(got.forcing["netFlux"], want.forcing["netFlux"])
    assert np.array_equal(got.forcing["surface_pH"], want.forcing["surface_pH"])
    ctx2.close()
    # and the reference, three steps
    ref = cols.copy()
    for _ in range(3):
        o.BGC_SourceSink(po, ref, True, nthreads=o.max_threads())
        o.BGC_SurfaceFluxes(po, ref, nthreads=o.max_threads())
    parity.compare_bgc_source_sink(ref, want)
    assert parity.nerr(want.forcing["surface_pH"], ref.forcing["surface_pH"]) <= parity.TOL_SOLVER


def test_mpas_thickness_weighted_tendency():
    """bgc_layout_soa_to_mpas_weighted: tend_mpas(n,k,cell) += layerThickness(k,cell) * tend_soa(cell,k,slot[n]),
    the way MPAS-Ocean folds the BGC tendencies into its thickness-weighted tracer tendencies."""
    import torch
    nT, nL, nC, nS = 30, 9, 777, 30
    ctx, _ = _ctx(nL, nC)
    g = torch.Generator(device="cpu").manual_seed(3)
    tend_mpas = torch.rand((nC, nL, nT), generator=g, dtype=torch.float64).cuda()
    h = (torch.rand((nC, nL), generator=g, dtype=torch.float64) * 100.0 + 1.0).cuda()     # (cell,k): k fastest
    soa = torch.rand((nS, nL, nC), generator=g, dtype=torch.float64).cuda()
    slot = [int(x) + 1 for x in torch.randperm(nS, generator=g)]
    want = tend_mpas.clone()
    for n in range(nT):
        want[:, :, n] = want[:, :, n] + (h * soa[slot[n] - 1].T)
    got = tend_mpas.clone()
    torch.cuda.synchronize()
    ctx.soa_to_mpas(soa.data_ptr(), got.data_ptr(), slot, nL, nC, alpha=1.0, beta=1.0, dev_weight=h.data_ptr())
    ctx.synchronize()
    assert (got - want).abs().max().item() <= 1e-13 * want.abs().max().item()
    # alpha scales the weighted term; weight = None is the unweighted form
    got2 = tend_mpas.clone()
    ctx.soa_to_mpas(soa.data_ptr(), got2.data_ptr(), slot, nL, nC, alpha=0.5, beta=0.0, dev_weight=h.data_ptr())
    ctx.synchronize()
    for n in range(nT):
        assert torch.allclose(got2[:, :, n], 0.5 * h * soa[slot[n] - 1].T, rtol=1e-15, atol=0)
    ctx.close()


def test_flavours_stay_separate_when_one_is_loaded_globally(o):
    """Both flavours export the same names.  With the production library in the process's GLOBAL symbol
    scope (what tests/test_fortran_shim.py does, and what a Fortran host's link line amounts to) the strict
    library's internal calls must still reach its OWN kernels (-Bsymbolic): its answer differs from the
    production build's in the last bits and is the one that sits within 1e-12 of the oracle."""
    import ctypes
    import os
    ctypes.CDLL(os.path.join(host.CSRC, host.LIB_NAME["prod"]), mode=ctypes.RTLD_GLOBAL)
    nL, nC = 40, 192
    po = o.Parms()
    res = {}
    for flavour in ("prod", "strict"):
        ctx, parms = _ctx(nL, nC, flavour=flavour)
        cols, _, _ = parity.make_bgc(nL, nC, parms, ragged=True)
        if flavour == "prod":
            ref = cols.copy()
            o.BGC_SourceSink(po, ref, True, nthreads=o.max_threads())
        res[flavour] = parity.run_gpu_bgc(ctx, cols, device_mode=True)
        ctx.close()
    assert not np.array_equal(res["prod"].BGC_tendencies, res["strict"].BGC_tendencies)
    worst = max(parity.nerr(res["strict"].BGC_tendencies[:, :, n], ref.BGC_tendencies[:, :, n]) for n in range(30))
    assert worst <= 1e-12, worst
