"""The carbonate DEVICE functions (ocean-bgc_b200/csrc/bgc_co2.cuh: co3_coeffs, solve_htotal,
co3_sat_vals - the code the CUDA kernels inline) compiled for the host (tests/host_twin/, one lane
standing in for a warp; test infrastructure) and compared with the translated reference on the CPU.
Both flavours of the product build are covered: production (custom exp/log, reciprocal + Newton
division, one exponential per pressure-corrected constant) and strict (-DBGC_STRICT: IEEE division,
library exp/log/pow).  This is not a substitute for the GPU parity tests - the kernels around
these functions only run on the device - but it shows at the ulp level what the device arithmetic
does to every equilibrium constant and to the solver's root, without a GPU.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import parity

sys.path.insert(0, os.path.join(parity.REPO, "oracle"))
import ref_translated as rt   # noqa: E402  (test infrastructure only)

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not rt.available(), reason="oracle/_ref/libbgc_ref.so not built")
NAMES = ["k1", "k2", "ff", "kw", "kb", "ks", "kf", "k1p", "k2p", "k3p", "ksi", "bt", "st", "ft"]


@pytest.fixture(scope="module", params=["production", "strict"])
def twin(request, tmp_path_factory):
    so = str(tmp_path_factory.mktemp("twin") / ("libco2_twin_%s.so" % request.param))
    flags = ["-DBGC_STRICT=1"] if request.param == "strict" else []
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-D_GNU_SOURCE", *flags,
                           "-I" + os.path.join(HERE, "host_twin", "stub"),
                           "-I" + os.path.join(parity.REPO, "ocean-bgc_b200", "csrc"),
                           "-o", so, os.path.join(HERE, "host_twin", "co2_twin.cpp")])
    return request.param, C.CDLL(so)


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


@pytest.fixture(scope="module")
def points():
    rng = np.random.default_rng(12)
    n = 3000
    p = dict(k=rng.choice([1, 2, 30, 60], n).astype(np.int32), depth=rng.uniform(0, 5500, n),
             temp=rng.uniform(-1.8, 31, n), salt=rng.uniform(0.5, 40, n), dic=rng.uniform(1800, 2400, n))
    p["ta"] = p["dic"] + rng.uniform(80, 420, n)
    p["pt"], p["sit"] = rng.uniform(0, 3, n), rng.uniform(0, 150, n)
    return p


def test_equilibrium_constants(twin, points):
    flavour, L = twin
    rt.RefParms()
    p, n = points, len(points["k"])
    out = np.zeros((n, 14))
    L.twin_co3_coeffs(C.c_int(n), ip(p["k"]), dp(p["depth"]), dp(p["temp"]), dp(p["salt"]), dp(out))
    ref = np.zeros((n, 14))
    for i in range(n):
        _, b = rt.call("co2calc__comp_co3_coeffs", int(p["k"][i]), p["depth"][i], p["temp"][i], p["salt"][i],
                       0.0, 0.0, 0.0, 0.0, 1)
        ref[i, :3] = b[5].value, b[6].value, b[7].value
        ref[i, 3:] = [rt.var("co2calc__" + nm)[0] for nm in NAMES[3:]]
    rel = np.abs(out - ref) / np.abs(ref)
    worst = {nm: float(rel[:, j].max()) for j, nm in enumerate(NAMES)}
    # exp() of arguments up to ~45 amplifies an argument error of one ulp to ~45 ulp of the result
    # Each constant is exp(sum of terms of size 30..150 that cancel to -5..-45): an error of a few ulp
    # of the largest term becomes a relative error of ~1e-13 of the constant.  Measured: production
    # 2.3e-13 (kb), strict 5e-15; the parity bound of the path is 1e-10.
    lim = 5e-13 if flavour == "production" else 5e-15
    assert max(worst.values()) <= lim, (flavour, worst)


def test_solver_root_and_saturation(twin, points):
    flavour, L = twin
    rt.RefParms()
    p, n = points, len(points["k"])
    for cold in (True, False):
        if cold:
            lo, hi = np.full(n, 6.0), np.full(n, 9.0)
        else:
            lo, hi = ph - 0.2, ph + 0.2
        h, st = np.zeros(n), np.zeros(n, dtype=np.int32)
        L.twin_htotal(C.c_int(n), ip(p["k"]), *[dp(p[k]) for k in ("depth", "temp", "salt", "dic", "ta", "pt", "sit")],
                      dp(lo), dp(hi), dp(h), ip(st))
        ph = np.array([rt.comp_CO3terms(int(p["k"][i]), *[float(p[k][i]) for k in
                                                          ("depth", "temp", "salt", "dic", "ta", "pt", "sit")],
                                        float(lo[i]), float(hi[i]))["pH"] for i in range(n)])
        href = 10.0 ** (-ph)
        assert not st.any()
        # both stop when |dx| < xacc = 1e-10 mol/kg (co2calc.F90:53); same trajectory -> far closer
        assert np.abs(h - href).max() <= 1e-10
        assert (np.abs(h - href) / href).max() <= (1e-9 if flavour == "production" else 1e-11), flavour
    calc, arag = np.zeros(n), np.zeros(n)
    L.twin_sat_vals(C.c_int(n), ip(p["k"]), dp(p["depth"]), dp(p["temp"]), dp(p["salt"]), dp(calc), dp(arag))
    ref = np.array([rt.comp_co3_sat_vals(int(p["k"][i]), p["depth"][i], p["temp"][i], p["salt"][i]) for i in range(n)])
    lim = 5e-13 if flavour == "production" else 2e-14     # measured 1.3e-13 / < 1e-14
    assert (np.abs(calc - ref[:, 0]) / ref[:, 0]).max() <= lim
    assert (np.abs(arag - ref[:, 1]) / ref[:, 1]).max() <= lim
