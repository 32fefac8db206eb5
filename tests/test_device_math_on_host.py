"""Accuracy of the device math (ocean-bgc_b200/csrc/bgc_math.cuh, production flavour) measured on
the CPU: the header is compiled for the host through a small stand-in for <cuda_runtime.h>
(tests/host_twin/, test infrastructure) - its own code, FMAs included; only the MUFU.RCP64H seed is
replaced by the worst seed its measured 2^-20 bound allows - and compared with 80-bit long-double
references over the argument ranges the kernels use.  Checks the ulp bounds stated in the header:
frcp <= 1.5, bexp <= 2.2 (2.15 measured), table-driven exp <= 1.3 (1.28), blog <= 3 (2.4).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)

pytestmark = pytest.mark.skipif(np.finfo(np.longdouble).nmant < 63, reason="needs 80-bit long double")


@pytest.fixture(scope="module")
def twin(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("twin") / "libmath_twin.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                           "-I" + os.path.join(HERE, "host_twin", "stub"),
                           "-I" + os.path.join(REPO, "ocean-bgc_b200", "csrc"),
                           "-o", so, os.path.join(HERE, "host_twin", "math_twin.cpp")])
    return C.CDLL(so)


def _call(fn, *xs):
    xs = [np.ascontiguousarray(x, dtype=np.float64) for x in xs]
    y = np.empty_like(xs[0])
    fn(C.c_int(len(y)), *[x.ctypes.data_as(C.POINTER(C.c_double)) for x in xs], y.ctypes.data_as(C.POINTER(C.c_double)))
    return y


def ulp_error(got, exact):
    """|got - exact| in units of the last place of the correctly rounded double"""
    ref = exact.astype(np.float64)
    ulp = np.spacing(np.abs(ref)).astype(np.longdouble)
    return np.abs(got.astype(np.longdouble) - exact) / ulp


def test_reciprocal(twin):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(1.0, 2.0, 400000), 10.0 ** rng.uniform(-30, 30, 400000),
                        -10.0 ** rng.uniform(-10, 10, 100000)])
    e = ulp_error(_call(twin.twin_rcp, x), np.longdouble(1.0) / x.astype(np.longdouble))
    assert e.max() <= 1.5, e.max()


def test_exp(twin):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-708.0, 709.0, 600000), rng.uniform(-40.0, 40.0, 600000),
                        rng.uniform(-1.0, 1.0, 300000), rng.normal(0, 1e-8, 1000)])
    exact = np.exp(x.astype(np.longdouble))
    assert ulp_error(_call(twin.twin_exp, x), exact).max() <= 2.2
    assert ulp_error(_call(twin.twin_exp_table, x), exact).max() <= 1.3
    below = np.array([-708.0001, -720.0, -745.0, -1e4])
    assert np.all(_call(twin.twin_exp, below) == 0.0) and np.all(_call(twin.twin_exp_table, below) == 0.0)
    assert _call(twin.twin_exp, np.array([0.0]))[0] == 1.0


def test_log(twin):
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(0.5, 2.0, 600000), 10.0 ** rng.uniform(-300, 300, 300000),
                        rng.uniform(250.0, 320.0, 300000),             # temperatures in kelvin
                        1.0 + rng.normal(0, 1e-6, 50000)])               # the cancellation region around 1
    x = x[x > 0]
    exact = np.log(x.astype(np.longdouble))
    e = ulp_error(_call(twin.twin_log, x), exact)
    far = np.abs(x - 1.0) > 1e-3
    assert e[far].max() <= 3.0, e[far].max()
    # next to 1 the result itself is tiny; the absolute error stays below 3 ulp of |x - 1|
    near = ~far
    err = np.abs(_call(twin.twin_log, x)[near].astype(np.longdouble) - exact[near])
    assert np.all(err <= 3 * np.spacing(np.abs(x[near] - 1.0) + np.finfo(float).tiny))


def test_pow_as_exp_log(twin):
    """x**y = exp(y log x): the error of log x is amplified by |y log x|; on this path
    (Q10 factors, Chl**0.4562, 0.99**(O2 - NO3), 10**(-pH)) that product stays below ~40."""
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(0.02, 50.0, 300000), np.full(100000, 10.0), np.full(100000, 0.99)])
    y = np.concatenate([rng.uniform(0.3, 2.0, 300000), -rng.uniform(4.0, 10.0, 100000), rng.uniform(-300, 300, 100000)])
    exact = np.exp(y.astype(np.longdouble) * np.log(x.astype(np.longdouble)))
    rel = np.abs(_call(twin.twin_pow, x, y).astype(np.longdouble) - exact) / exact
    assert rel.max() <= 2e-14, float(rel.max())      # four orders below the 1e-10 parity bound
