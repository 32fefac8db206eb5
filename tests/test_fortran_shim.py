"""The drop-in Fortran shim (ocean-bgc_b200/fortran/: modules BGC_mod, DMS_mod, MACROS_mod with the
reference's procedure names, argument lists and derived types, forwarding to the C ABI through
ISO_C_BINDING) EXECUTED, not just written.

The image has no Fortran compiler, so the shim is translated to C by oracle/f90c.py together with
the reference's own unchanged BGC_parms / DMS_parms / MACROS_parms (the module set its Makefile
compiles) -> oracle/_ref/libbgc_shim.so.  The caller side below does what MPAS does: it fills the
reference's derived types (allocatable components, level index fastest), runs the reference's
*_parms_init, the shim's *_init, and then calls BGC_SourceSink & co with the reference signatures.
Underneath, the C-ABI symbols resolve to

  * on CPU (this file's non-gpu tests): tests/mock_abi/mock_bgc_b200.c, a stand-in backed by the
    oracle - the answers must equal the oracle's bit for bit, which checks every component ->
    struct-member assignment, the extents, the flags, the parameter flattening and the ctx
    life cycle of the shim;
  * on the GPU box (`-m gpu`): the real libbgc_b200.so - the complete drop-in chain
    reference-typed caller -> shim -> C ABI -> CUDA, compared with the translated reference.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import parity

pkg = parity.pkg
abi = pkg.abi
o = parity.oracle()
sys.path.insert(0, os.path.join(parity.REPO, "oracle"))
import ref_translated as rt   # noqa: E402  (test infrastructure only)

SHIM = os.path.join(rt.REFDIR, "libbgc_shim.so")
SHIM_META = os.path.join(rt.REFDIR, "shim_meta.json")
MOCK = os.path.join(rt.REFDIR, "libmock_bgc_b200.so")
REAL = os.path.join(parity.REPO, "ocean-bgc_b200", "csrc", "libbgc_b200.so")

if not os.path.exists(SHIM) and rt.can_build():
    rt.build()
pytestmark = pytest.mark.skipif(not (os.path.exists(SHIM) and os.path.exists(SHIM_META)),
                                reason="oracle/_ref/libbgc_shim.so not built (needs the reference sources once)")


def _same(a, b, what):
    if not np.array_equal(a, b, equal_nan=True):
        raise AssertionError("%s differs: max |d| = %.3e" % (what, float(np.nanmax(np.abs(a - b)))))


def run_through_shim(L, po, cols, dms, mac, passes=2):
    """What an MPAS-like caller does, through the reference API names."""
    rp = rt.RefParms(po, L=L)              # BGC_parms_init, BGC_init (shim), DMS_*, MACROS_*
    for _ in range(passes):
        rt.BGC_SourceSink(rp, cols, True)  # shim BGC_mod::BGC_SourceSink
    rt.BGC_SurfaceFluxes(rp, cols)
    rt.DMS_SourceSink(rp, dms)
    rt.DMS_SurfaceFluxes(rp, dms)
    rt.MACROS_SourceSink(rp, mac)
    return rp


def test_shim_marshalling_against_the_oracle():
    L = rt.TLib(SHIM, SHIM_META, preload=[MOCK])
    mock = C.CDLL(MOCK, mode=C.RTLD_GLOBAL)
    po = o.Parms()
    nL, nC, nCols = 24, 96, 90
    cols, dms, mac = parity.make_bgc(nL, nC, po, ragged=True, nColumns=nCols, with_dms=True, with_macros=True)
    for c in (cols, dms, mac):
        parity.poison_outputs(c)
    a, da, ma = cols.copy(), dms.copy(), mac.copy()
    o.BGC_SourceSink(po, a, True); o.BGC_SourceSink(po, a, True); o.BGC_SurfaceFluxes(po, a)
    o.DMS_SourceSink(po, da); o.DMS_SurfaceFluxes(po, da); o.MACROS_SourceSink(po, ma)
    b, db, mb = cols.copy(), dms.copy(), mac.copy()
    calls0 = mock.mock_compute_calls()
    rp = run_through_shim(L, po, b, db, mb)
    assert mock.mock_compute_calls() - calls0 == 6
    _same(b.BGC_tendencies, a.BGC_tendencies, "BGC tendencies")
    _same(b.PH_PREV_3D, a.PH_PREV_3D, "PH_PREV_3D")
    _same(b.PH_PREV_ALT_CO2_3D, a.PH_PREV_ALT_CO2_3D, "PH_PREV_ALT_CO2_3D")
    for n in a.diag:
        _same(b.diag[n], a.diag[n], n)
    for n in a.forcing:
        _same(b.forcing[n], a.forcing[n], "forcing " + n)
    for n in a.flux_diag:
        _same(b.flux_diag[n], a.flux_diag[n], "flux diag " + n)
    _same(db.DMS_tendencies, da.DMS_tendencies, "DMS tendencies")
    for n in da.diag:
        _same(db.diag[n], da.diag[n], "DMS " + n)
    for n in da.flux_diag:
        _same(db.flux_diag[n], da.flux_diag[n], "DMS flux diag " + n)
    _same(db.forcing["netFlux"], da.forcing["netFlux"], "DMS netFlux")
    _same(mb.MACROS_tendencies, ma.MACROS_tendencies, "MACROS tendencies")
    for n in ma.diag:
        _same(mb.diag[n], ma.diag[n], "MACROS " + n)
    # BGC_init of the shim: same index wiring and names as the reference's
    ref = rt.RefParms(po)
    for i in range(4):
        for f in ("chl_ind", "c_ind", "fe_ind", "si_ind", "caco3_ind"):
            assert getattr(rp.autotrophs[i], f) == getattr(ref.autotrophs[i], f), (i, f)
    for which, n in (("ind", 30), ("dms_ind", 14), ("macros_ind", 8)):
        for i in range(n):
            assert rp.name(which, "short_name", i) == ref.name(which, "short_name", i), (which, i)
            assert rp.name(which, "units", i) == ref.name(which, "units", i), (which, i)


def test_shim_context_life_cycle_and_parameter_changes():
    """One ctx per (thread of the) host, re-created only when a larger block shows up; parameter
    changes made through the reference's module variables reach the library on the next call."""
    L = rt.TLib(SHIM, SHIM_META, preload=[MOCK])
    mock = C.CDLL(MOCK, mode=C.RTLD_GLOBAL)
    po = o.Parms()
    rp = rt.RefParms(po, L=L)
    L.call("bgc_b200_runtime__bgc_b200_finalize")      # whatever an earlier test left behind
    assert mock.mock_live_contexts() == 0
    created0 = mock.mock_created_contexts()
    small, _, _ = parity.make_bgc(10, 16, po)
    rt.BGC_SourceSink(rp, small, True)
    rt.BGC_SourceSink(rp, small, True)
    n1 = mock.mock_created_contexts() - created0
    big, _, _ = parity.make_bgc(20, 64, po)
    rt.BGC_SourceSink(rp, big, True)
    n2 = mock.mock_created_contexts() - created0
    rt.BGC_SourceSink(rp, small, True)           # fits the larger ctx: no new one
    assert (n1, n2, mock.mock_created_contexts() - created0) == (1, 2, 2)
    assert mock.mock_live_contexts() == 1
    assert L.var("bgc_b200_runtime__ctx_levels").value == 20
    assert L.var("bgc_b200_runtime__ctx_columns").value == 64
    # a namelist-style change of a module variable
    po2 = o.Parms()
    po2.bgc.parm_labile_ratio = 0.6
    po2.bgc.parm_o2_min = 7.0
    rp.sync_from(po2)
    a, b = big.copy(), big.copy()
    o.BGC_SourceSink(po2, a, True)
    rt.BGC_SourceSink(rp, b, True)
    _same(b.BGC_tendencies, a.BGC_tendencies, "tendencies after a parameter change")
    L.call("bgc_b200_runtime__bgc_b200_finalize")
    assert mock.mock_live_contexts() == 0


def test_shim_device_selection(monkeypatch):
    """bgc_b200_runtime: the device comes from bgc_b200_set_device, else from BGC_B200_DEVICE, else 0
    (one MPI rank per GPU passes its node-local rank).  Module state is per thread in the translation,
    so every case runs in a thread of its own, as a fresh process would."""
    import threading
    L = rt.TLib(SHIM, SHIM_META, preload=[MOCK])
    mock = C.CDLL(MOCK, mode=C.RTLD_GLOBAL)
    po = o.Parms()
    seen = {}

    def case(name, env, set_device):
        def run():
            if env is None:
                os.environ.pop("BGC_B200_DEVICE", None)
            else:
                os.environ["BGC_B200_DEVICE"] = env
            rp = rt.RefParms(po, L=L)
            if set_device is not None:
                L.call("bgc_b200_runtime__bgc_b200_set_device", set_device)
            cols, _, _ = parity.make_bgc(5, 8, po)
            rt.BGC_SourceSink(rp, cols, True)
            seen[name] = mock.mock_last_device()
            L.call("bgc_b200_runtime__bgc_b200_finalize")
        t = threading.Thread(target=run)
        t.start(); t.join()

    keep = os.environ.get("BGC_B200_DEVICE")
    try:
        case("default", None, None)
        case("env", "3", None)
        case("env garbage", "gpu-one", None)
        case("explicit beats env", "3", 5)
    finally:
        if keep is None:
            os.environ.pop("BGC_B200_DEVICE", None)
        else:
            os.environ["BGC_B200_DEVICE"] = keep
    assert seen == {"default": 0, "env": 3, "env garbage": 0, "explicit beats env": 5}, seen


def test_shim_fails_loudly(tmp_path):
    """The reference has no error reporting; the shim must not swallow a failed GPU call (there is no
    CPU fallback): message of the library on stderr, then `error stop`.  Run in a child process."""
    import subprocess
    prog = (
        "import sys, os\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import parity, ref_translated as rt\n"
        "o = parity.oracle(); po = o.Parms()\n"
        "L = rt.TLib(%r, %r, preload=[%r])\n"
        "rp = rt.RefParms(po, L=L)\n"
        "cols, _, _ = parity.make_bgc(5, 8, po)\n"
        "rt.BGC_SourceSink(rp, cols, True)\n"
        "print('survived')\n") % (parity.HERE, os.path.join(parity.REPO, "oracle"), SHIM, SHIM_META, MOCK)
    env = dict(os.environ, MOCK_BGC_FAIL="1")
    r = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode != 0 and "survived" not in r.stdout
    assert "bgc_source_sink failed with code" in r.stderr and "-2" in r.stderr
    assert "simulated CUDA failure" in r.stderr and "GPU hot path failed" in r.stderr
    ok = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, timeout=300)
    assert ok.returncode == 0 and "survived" in ok.stdout


def test_shim_forwards_every_parameter():
    """push_params / the DMS and MACROS flattening list every field by hand: perturb ALL tunables,
    the whole functional-group table and the restoring switches at random and demand the oracle's
    bits - a field forwarded to the wrong member (or not at all) changes the answer."""
    sys.path.insert(0, os.path.join(parity.REPO, "scripts"))
    import fuzz_oracle_vs_reference as gen
    L = rt.TLib(SHIM, SHIM_META, preload=[MOCK])
    for seed in (1, 2, 4, 5):
        rng = np.random.default_rng(seed)
        po = o.Parms()
        gen.perturb_parms(po, rng)
        rp = rt.RefParms(po, L=L).sync_from(po)
        cols, dms, mac = parity.make_bgc(17, 40, po, ragged=True, with_dms=True, with_macros=True, seed=77 + seed)
        cols.forcing["NUTR_RESTORE_RTAU"][...] = rng.uniform(0.0, 1e-6, cols.forcing["NUTR_RESTORE_RTAU"].shape)
        for nm, slot in (("NO3_CLIM", po.ind.no3_ind), ("PO4_CLIM", po.ind.po4_ind), ("SiO3_CLIM", po.ind.sio3_ind)):
            cols.forcing[nm][...] = cols.BGC_tracers[:, :, slot - 1] * 1.1
        a, da, ma = cols.copy(), dms.copy(), mac.copy()
        o.BGC_SourceSink(po, a, True); o.BGC_SurfaceFluxes(po, a)
        o.DMS_SourceSink(po, da); o.DMS_SurfaceFluxes(po, da); o.MACROS_SourceSink(po, ma)
        b, db, mb = cols.copy(), dms.copy(), mac.copy()
        rt.BGC_SourceSink(rp, b, True); rt.BGC_SurfaceFluxes(rp, b)
        rt.DMS_SourceSink(rp, db); rt.DMS_SurfaceFluxes(rp, db); rt.MACROS_SourceSink(rp, mb)
        _same(b.BGC_tendencies, a.BGC_tendencies, "seed %d BGC tendencies" % seed)
        for n in a.diag:
            _same(b.diag[n], a.diag[n], "seed %d %s" % (seed, n))
        _same(b.forcing["netFlux"], a.forcing["netFlux"], "seed %d netFlux" % seed)
        _same(db.DMS_tendencies, da.DMS_tendencies, "seed %d DMS tendencies" % seed)
        _same(db.forcing["netFlux"], da.forcing["netFlux"], "seed %d DMS netFlux" % seed)
        _same(mb.MACROS_tendencies, ma.MACROS_tendencies, "seed %d MACROS tendencies" % seed)
    L.call("bgc_b200_runtime__bgc_b200_finalize")


@pytest.mark.gpu
def test_shim_drop_in_chain_on_the_gpu():
    """reference-typed caller -> shim -> C ABI -> CUDA, against the translated reference."""
    if "libmock_bgc_b200" in open("/proc/self/maps").read():
        pytest.skip("the CPU mock of the C ABI is loaded in this process (run with -m gpu)")
    L = rt.TLib(SHIM, SHIM_META, preload=[REAL])
    po = pkg.host.Parms()
    nL, nC, nCols = 40, 258, 250
    cols, dms, mac = parity.make_bgc(nL, nC, po, ragged=True, nColumns=nCols, with_dms=True, with_macros=True)
    for c in (cols, dms, mac):
        parity.poison_outputs(c)
    rr = rt.RefParms(po)
    a, da, ma = cols.copy(), dms.copy(), mac.copy()
    rt.BGC_SourceSink(rr, a, True)
    b, db, mb = cols.copy(), dms.copy(), mac.copy()
    rp = rt.RefParms(po, L=L)
    rt.BGC_SourceSink(rp, b, True)
    parity.compare_bgc_source_sink(a, b)              # cold pass
    rt.BGC_SourceSink(rr, a, True); rt.BGC_SurfaceFluxes(rr, a)
    rt.BGC_SourceSink(rp, b, True); rt.BGC_SurfaceFluxes(rp, b)
    parity.compare_bgc_source_sink(a, b)              # warm pass
    cm = np.arange(nC) < nCols
    assert parity.nerr(b.forcing["netFlux"][cm], a.forcing["netFlux"][cm]) <= parity.TOL_SOLVER
    rt.DMS_SourceSink(rr, da); rt.DMS_SurfaceFluxes(rr, da); rt.MACROS_SourceSink(rr, ma)
    rt.DMS_SourceSink(rp, db); rt.DMS_SurfaceFluxes(rp, db); rt.MACROS_SourceSink(rp, mb)
    assert parity.nerr(db.DMS_tendencies, da.DMS_tendencies) <= parity.TOL_TEND
    assert parity.nerr(mb.MACROS_tendencies, ma.MACROS_tendencies) <= parity.TOL_TEND
    assert parity.nerr(db.forcing["netFlux"][cm], da.forcing["netFlux"][cm]) <= parity.TOL_TEND
    L.call("bgc_b200_runtime__bgc_b200_finalize")


def _co2calc_trio_through(L, pts, k, n):
    """co2calc_1point, comp_CO3terms, comp_co3_sat_vals of module co2calc with the reference's scalar
    argument lists (co2calc.F90:75, :214, :1096), point by point."""
    out = {nm: np.zeros(n) for nm in ("ph1", "co2star", "dco2star", "pco2surf", "dpco2", "ph", "h2co3", "hco3",
                                      "co3", "sat_calc", "sat_arag")}
    for i in range(n):
        a = {nm: float(pts[nm][i]) for nm in pts}
        _, b = L.call("co2calc__co2calc_1point", a["depth"], 1, 1, a["temp"], a["salt"], a["dic"], a["ta"], a["pt"],
                      a["sit"], a["phlo"], a["phhi"], 0.0, a["xco2"], a["atmpres"], 0.0, 0.0, 0.0, 0.0)
        out["ph1"][i], out["co2star"][i], out["dco2star"][i] = b[11].value, b[14].value, b[15].value
        out["pco2surf"][i], out["dpco2"][i] = b[16].value, b[17].value
        _, b = L.call("co2calc__comp_co3terms", int(k[i]), a["depth"], 1, a["temp"], a["salt"], a["dic"], a["ta"],
                      a["pt"], a["sit"], a["phlo"], a["phhi"], 0.0, 0.0, 0.0, 0.0)
        assert (b[9].value, b[10].value) == (a["phlo"], a["phhi"])      # the brackets come back unchanged
        out["ph"][i], out["h2co3"][i], out["hco3"][i], out["co3"][i] = (x.value for x in b[11:15])
        _, b = L.call("co2calc__comp_co3_sat_vals", int(k[i]), a["depth"], a["temp"], a["salt"], 0.0, 0.0)
        out["sat_calc"][i], out["sat_arag"][i] = b[4].value, b[5].value
    return out


def _co2calc_trio_reference(pts, k, n):
    out = {nm: np.zeros(n) for nm in ("ph1", "co2star", "dco2star", "pco2surf", "dpco2", "ph", "h2co3", "hco3",
                                      "co3", "sat_calc", "sat_arag")}
    for i in range(n):
        a = {nm: float(pts[nm][i]) for nm in pts}
        r = rt.co2calc_1point(*[a[nm] for nm in ("depth", "temp", "salt", "dic", "ta", "pt", "sit", "phlo", "phhi",
                                                  "xco2", "atmpres")])
        out["ph1"][i], out["co2star"][i], out["dco2star"][i] = r["ph"], r["co2star"], r["dco2star"]
        out["pco2surf"][i], out["dpco2"][i] = r["pco2surf"], r["dpco2"]
        r = rt.comp_CO3terms(int(k[i]), a["depth"], a["temp"], a["salt"], a["dic"], a["ta"], a["pt"], a["sit"],
                             a["phlo"], a["phhi"])
        out["ph"][i], out["h2co3"][i], out["hco3"][i], out["co3"][i] = r["pH"], r["H2CO3"], r["HCO3"], r["CO3"]
        out["sat_calc"][i], out["sat_arag"][i] = rt.comp_co3_sat_vals(int(k[i]), a["depth"], a["temp"], a["salt"])
    return out


def _trio_points(n):
    pts = pkg.synth_co2_points(n)
    rng = np.random.default_rng(5)
    k = rng.integers(1, 61, n).astype(np.int32)
    k[:4] = 1
    pts["depth"] = np.where(k == 1, 5.0, rng.uniform(10.0, 5500.0, n))
    return pts, k


def test_shim_co2calc_module_against_the_reference():
    """module co2calc of the shim (the reference's public trio, co2calc.F90:24) over the mock C ABI:
    the reference's bits (the mock is the oracle, and oracle == translated reference)."""
    L = rt.TLib(SHIM, SHIM_META, preload=[MOCK])
    C.CDLL(MOCK, mode=C.RTLD_GLOBAL)
    n = 48
    pts, k = _trio_points(n)
    got, ref = _co2calc_trio_through(L, pts, k, n), _co2calc_trio_reference(pts, k, n)
    for nm in ref:
        _same(got[nm], ref[nm], "co2calc module: " + nm)
    L.call("bgc_b200_runtime__bgc_b200_finalize")


@pytest.mark.gpu
def test_shim_co2calc_module_on_the_gpu():
    """the same scalar calls, reference-typed caller -> shim module co2calc -> C ABI -> CUDA"""
    if "libmock_bgc_b200" in open("/proc/self/maps").read():
        pytest.skip("the CPU mock of the C ABI is loaded in this process (run with -m gpu)")
    L = rt.TLib(SHIM, SHIM_META, preload=[REAL])
    n = 24
    pts, k = _trio_points(n)
    got, ref = _co2calc_trio_through(L, pts, k, n), _co2calc_trio_reference(pts, k, n)
    for nm in ref:
        e = parity.nerr(got[nm], ref[nm])
        assert e <= parity.TOL_SOLVER, (nm, e)
    L.call("bgc_b200_runtime__bgc_b200_finalize")
