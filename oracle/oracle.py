"""ctypes binding of the CPU oracle (oracle/libbgc_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Pinned bit for bit against
the reference's own sources machine-translated to C (oracle/f90c.py ->
oracle/_ref/libbgc_ref.so, oracle/ref_translated.py); see bgc_oracle.h for the
caveat (a translation, not a gfortran build).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
if REPO not in sys.path:
    sys.path.insert(0, REPO)
import __graft_entry__ as _ge  # noqa: E402

pkg = _ge.load_package()
abi = pkg.abi

LIB_PATH = os.path.join(HERE, "libbgc_oracle.so")


class Co2Save(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("kw", "kb", "ks", "kf", "k1p", "k2p", "k3p", "ksi", "bt", "st", "ft",
                 "dic", "ta", "pt", "sit")]


class SolverStats(C.Structure):
    _fields_ = [("talk_row_calls", C.c_long), ("bracket_grow", C.c_long),
                ("newton_iters", C.c_long), ("no_convergence", C.c_long)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_O2SAT_singleValue.restype = C.c_double
        _lib.oracle_O2SAT_singleValue.argtypes = [C.c_double] * 3
        for f in ("oracle_SCHMIDT_O2_singleValue", "oracle_SCHMIDT_CO2_singleValue",
                  "oracle_SCHMIDT_DMS_singleValue"):
            getattr(_lib, f).restype = C.c_double
            getattr(_lib, f).argtypes = [C.c_double]
        _lib.oracle_dust_to_Fe.restype = C.c_double
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return int(lib().oracle_max_threads())


class Parms:
    """Default parameter tables as the reference initialisers produce them."""

    def __init__(self, default_real_8=False):
        self.bgc = abi.BgcParams()
        self.autotrophs = abi.BgcAutotroph4()
        self.ind = abi.BgcIndices()
        self.dms = abi.DmsParams()
        self.dms_ind = abi.DmsIndices()
        self.macros = abi.MacrosParams()
        self.macros_ind = abi.MacrosIndices()
        # host-chosen tracer slots: declaration order 1..N
        for i, (n, _) in enumerate(abi.BgcIndices._fields_[:abi.BGC_TRACER_CNT]):
            setattr(self.ind, n, i + 1)
        for i, (n, _) in enumerate(abi.DmsIndices._fields_):
            setattr(self.dms_ind, n, i + 1)
        for i, (n, _) in enumerate(abi.MacrosIndices._fields_):
            setattr(self.macros_ind, n, i + 1)
        L = lib()
        L.oracle_BGC_parms_init(C.byref(self.bgc), self.autotrophs, C.byref(self.ind),
                                C.c_int(int(default_real_8)))
        L.oracle_BGC_init(C.byref(self.ind), self.autotrophs)
        L.oracle_DMS_parms_init(C.byref(self.dms))
        L.oracle_MACROS_parms_init(C.byref(self.macros))

    def permute_tracers(self, perm):
        """Re-wire the 30 BGC tracer slots (perm: 0-based permutation) — the
        host, not the library, chooses the numeric indices (BGC_parms.F90:82-112)."""
        for i, (n, _) in enumerate(abi.BgcIndices._fields_[:abi.BGC_TRACER_CNT]):
            setattr(self.ind, n, int(perm[i]) + 1)
        lib().oracle_BGC_init(C.byref(self.ind), self.autotrophs)


def O2SAT(sst, sss, t0=273.15):
    return lib().oracle_O2SAT_singleValue(sst, sss, t0)


def co3_coeffs(k, depth, temp, salt):
    """dict of the equilibrium constants of comp_co3_coeffs for arrays of points."""
    k = np.ascontiguousarray(k, dtype=np.int32)
    depth, temp, salt = (np.ascontiguousarray(a, dtype=np.float64) for a in (depth, temp, salt))
    n = len(k)
    out = np.zeros((n, 14))
    lib().oracle_co3_coeffs_points(C.c_int(n), abi.iptr(k), abi.dptr(depth), abi.dptr(temp),
                                   abi.dptr(salt), abi.dptr(out))
    names = ["k0", "k1", "k2", "ff", "kw", "kb", "ks", "kf", "k1p", "k2p", "k3p", "ksi", "bt", "st"]
    return {nm: out[:, i].copy() for i, nm in enumerate(names)}


def co3_sat_vals(k, depth, temp, salt):
    a = C.c_double()
    b = C.c_double()
    lib().oracle_comp_co3_sat_vals(C.c_int(int(k)), C.c_double(depth), C.c_double(temp),
                                   C.c_double(salt), C.byref(a), C.byref(b))
    return a.value, b.value


def comp_CO3terms(k, depth, temp, salt, dic, ta, pt, sit, phlo, phhi):
    lo, hi = C.c_double(phlo), C.c_double(phhi)
    ph, h2, h1, c3 = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    st = SolverStats()
    lib().oracle_comp_CO3terms(C.c_int(int(k)), C.c_double(depth), C.c_int(1), C.c_double(temp),
                               C.c_double(salt), C.c_double(dic), C.c_double(ta), C.c_double(pt),
                               C.c_double(sit), C.byref(lo), C.byref(hi), C.byref(ph),
                               C.byref(h2), C.byref(h1), C.byref(c3), C.byref(st))
    return dict(pH=ph.value, H2CO3=h2.value, HCO3=h1.value, CO3=c3.value,
                talk_row_calls=st.talk_row_calls, bracket_grow=st.bracket_grow,
                newton_iters=st.newton_iters)


def co2calc_points(pts, nthreads=1):
    """Batched co2calc_1point.  pts: dict with depth,temp,salt,dic,ta,pt,sit,phlo,phhi,xco2,atmpres."""
    n = len(pts["temp"])
    a = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in pts.items()}
    out = {k: np.zeros(n) for k in ("ph", "co2star", "dco2star", "pco2surf", "dpco2")}
    st = SolverStats()
    lib().oracle_co2calc_points(
        C.c_int(n), *[abi.dptr(a[k]) for k in ("depth", "temp", "salt", "dic", "ta", "pt", "sit",
                                                "phlo", "phhi", "xco2", "atmpres")],
        *[abi.dptr(out[k]) for k in ("ph", "co2star", "dco2star", "pco2surf", "dpco2")],
        C.byref(st), C.c_int(nthreads))
    out["stats"] = dict(talk_row_calls=st.talk_row_calls, bracket_grow=st.bracket_grow,
                        newton_iters=st.newton_iters, no_convergence=st.no_convergence)
    return out


def BGC_SourceSink(parms, cols, alt_co2_use_eco=True, nthreads=1):
    """Reference-named entry: runs the oracle in place on a BgcColumns container."""
    cin, cfo, cout, cdg = cols.c_input(), cols.c_forcing(), cols.c_output(), cols.c_diag()
    st = SolverStats()
    lib().oracle_BGC_SourceSink(C.byref(parms.bgc), parms.autotrophs, C.byref(parms.ind),
                                C.byref(cin), C.byref(cfo), C.byref(cout), C.byref(cdg),
                                C.c_int(cols.nLevelsMax), C.c_int(cols.nColumnsMax),
                                C.c_int(cols.nColumns), C.c_int(int(alt_co2_use_eco)),
                                C.c_int(nthreads), C.byref(st))
    return dict(talk_row_calls=st.talk_row_calls, bracket_grow=st.bracket_grow,
                newton_iters=st.newton_iters, no_convergence=st.no_convergence)


def BGC_SurfaceFluxes(parms, cols, nthreads=1):
    cin, cfo, cfd = cols.c_input(), cols.c_forcing(), cols.c_flux_diag()
    lib().oracle_BGC_SurfaceFluxes(C.byref(parms.bgc), C.byref(parms.ind), C.byref(cin),
                                   C.byref(cfo), C.byref(cfd), C.c_int(cols.nLevelsMax),
                                   C.c_int(cols.nColumnsMax), C.c_int(cols.nColumns),
                                   C.c_int(nthreads))


def DMS_SourceSink(parms, cols, nthreads=1):
    cin, cfo, cout, cdg = cols.c_input(), cols.c_forcing(), cols.c_output(), cols.c_diag()
    lib().oracle_DMS_SourceSink(C.byref(parms.dms), C.byref(parms.dms_ind), C.byref(cin),
                                C.byref(cfo), C.byref(cout), C.byref(cdg),
                                C.c_int(cols.nLevelsMax), C.c_int(cols.nColumnsMax),
                                C.c_int(cols.nColumns), C.c_int(nthreads))


def DMS_SurfaceFluxes(parms, cols):
    cin, cfo, cfd = cols.c_input(), cols.c_forcing(), cols.c_flux_diag()
    lib().oracle_DMS_SurfaceFluxes(C.byref(parms.dms), C.byref(parms.dms_ind), C.byref(cin),
                                   C.byref(cfo), C.byref(cfd), C.c_int(cols.nLevelsMax),
                                   C.c_int(cols.nColumnsMax), C.c_int(cols.nColumns))


def MACROS_SourceSink(parms, cols, nthreads=1):
    cin, cout, cdg = cols.c_input(), cols.c_output(), cols.c_diag()
    lib().oracle_MACROS_SourceSink(C.byref(parms.macros), C.byref(parms.macros_ind),
                                   C.byref(cin), C.byref(cout), C.byref(cdg),
                                   C.c_int(cols.nLevelsMax), C.c_int(cols.nColumnsMax),
                                   C.c_int(cols.nColumns), C.c_int(nthreads))
