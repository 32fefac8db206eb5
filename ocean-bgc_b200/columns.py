"""Host-side containers mirroring the reference's derived types.

`BgcColumns`, `DmsColumns`, `MacrosColumns` hold the numpy arrays of
BGC_input/forcing/output/diagnostics_type (BGC_parms.F90:127-321) and their
DMS / MACROS analogues (DMS_parms.F90:85-154, MACROS_parms.F90:79-113) in the
reference's Fortran layout — A(k, column[, n]) with the level index fastest —
and build the C-ABI argument blocks (include/bgc_b200.h) that point at them.
"""
import ctypes as C

import numpy as np

from . import abi


def _f(shape, alloc=None, dtype=np.float64):
    """Fortran-layout array.  `alloc(nelem, dtype) -> 1-D ndarray` lets the caller
    supply the storage (e.g. page-locked memory for the C ABI's host path)."""
    if alloc is None:
        return np.zeros(shape, dtype=dtype, order="F")
    n = int(np.prod(shape))
    flat = alloc(n, dtype)
    flat[...] = 0
    return flat.reshape(shape, order="F")


class BgcColumns:
    """BGC_input_type + BGC_forcing_type + BGC_output_type + BGC_diagnostics_type
    + BGC_flux_diagnostics_type for (nLevelsMax, nColumnsMax)."""

    def __init__(self, nLevelsMax, nColumnsMax, nColumns=None, alloc=None, diagnostics=True):
        nL, nC = int(nLevelsMax), int(nColumnsMax)
        self.nLevelsMax, self.nColumnsMax = nL, nC
        self.nColumns = nC if nColumns is None else int(nColumns)
        _g = globals()["_f"]

        def _f(shape, dtype=np.float64):   # storage from `alloc` when given
            return _g(shape, alloc, dtype)
        # BGC_input_type
        self.BGC_tracers = _f((nL, nC, abi.BGC_TRACER_CNT))
        self.PotentialTemperature = _f((nL, nC))
        self.Salinity = _f((nL, nC))
        self.cell_center_depth = _f((nL, nC))
        self.cell_thickness = _f((nL, nC))
        self.cell_bottom_depth = _f((nL, nC))
        self.cell_latitude = _f((nC,))
        self.number_of_active_levels = _f((nC,), np.int32)
        # BGC_forcing_type
        self.forcing = {}
        for n in abi.BGC_FORCING_K2:
            self.forcing[n] = _f((nL, nC))
        for n in abi.BGC_FORCING_C1:
            self.forcing[n] = _f((nC,))
        for n in abi.BGC_FORCING_FLUX:
            self.forcing[n] = _f((nC, abi.BGC_TRACER_CNT))
        self.lcalc_O2_gas_flux = 1
        self.lcalc_CO2_gas_flux = 1
        # BGC_output_type
        self.BGC_tendencies = _f((nL, nC, abi.BGC_TRACER_CNT))
        self.PH_PREV_3D = _f((nL, nC))
        self.PH_PREV_ALT_CO2_3D = _f((nL, nC))
        # BGC_diagnostics_type
        self.diag = {}
        if diagnostics:
            for n in abi.BGC_DIAG_K2:
                self.diag[n] = _f((nL, nC))
            for n in abi.BGC_DIAG_KA:
                self.diag[n] = _f((nL, nC, abi.BGC_AUTOTROPH_CNT))
            for n in abi.BGC_DIAG_CA:
                self.diag[n] = _f((nC, abi.BGC_AUTOTROPH_CNT))
            for n in abi.BGC_DIAG_C1:
                self.diag[n] = _f((nC,))
        # BGC_flux_diagnostics_type
        self.flux_diag = {n: _f((nC,)) for n in abi.BGC_FLUX_DIAG}

    # ---- C-ABI argument blocks (keep `self` alive while they are in use)
    def c_input(self):
        s = abi.BgcInput()
        for n in ("BGC_tracers", "PotentialTemperature", "Salinity", "cell_center_depth",
                  "cell_thickness", "cell_bottom_depth", "cell_latitude"):
            setattr(s, n, abi.fptr(getattr(self, n)))
        s.number_of_active_levels = abi.iptr(self.number_of_active_levels)
        return s

    def c_forcing(self):
        s = abi.BgcForcing()
        for n, a in self.forcing.items():
            setattr(s, n, abi.fptr(a))
        s.lcalc_O2_gas_flux = int(self.lcalc_O2_gas_flux)
        s.lcalc_CO2_gas_flux = int(self.lcalc_CO2_gas_flux)
        return s

    def c_output(self):
        s = abi.BgcOutput()
        s.BGC_tendencies = abi.fptr(self.BGC_tendencies)
        s.PH_PREV_3D = abi.fptr(self.PH_PREV_3D)
        s.PH_PREV_ALT_CO2_3D = abi.fptr(self.PH_PREV_ALT_CO2_3D)
        return s

    def c_diag(self, enabled=True):
        s = abi.BgcDiagnostics()
        if enabled:
            for n, a in self.diag.items():
                setattr(s, n, abi.fptr(a))
        return s

    def c_flux_diag(self):
        s = abi.BgcFluxDiagnostics()
        for n, a in self.flux_diag.items():
            setattr(s, n, abi.fptr(a))
        return s

    def copy(self):
        o = BgcColumns(self.nLevelsMax, self.nColumnsMax, self.nColumns, diagnostics=bool(self.diag))
        for n in ("BGC_tracers", "PotentialTemperature", "Salinity", "cell_center_depth",
                  "cell_thickness", "cell_bottom_depth", "cell_latitude",
                  "number_of_active_levels", "BGC_tendencies", "PH_PREV_3D",
                  "PH_PREV_ALT_CO2_3D"):
            getattr(o, n)[...] = getattr(self, n)
        for d in ("forcing", "diag", "flux_diag"):
            for n, a in getattr(self, d).items():
                getattr(o, d)[n][...] = a
        o.lcalc_O2_gas_flux, o.lcalc_CO2_gas_flux = self.lcalc_O2_gas_flux, self.lcalc_CO2_gas_flux
        return o

    def active_mask(self):
        """(nLevelsMax, nColumnsMax) bool: cell is active (k <= kmax, column < nColumns)."""
        k = np.arange(1, self.nLevelsMax + 1)[:, None]
        kmax = self.number_of_active_levels.copy()
        kmax[self.nColumns:] = 0
        return k <= kmax[None, :]


class DmsColumns:
    def __init__(self, nLevelsMax, nColumnsMax, nColumns=None, alloc=None, diagnostics=True):
        nL, nC = int(nLevelsMax), int(nColumnsMax)
        self.nLevelsMax, self.nColumnsMax = nL, nC
        self.nColumns = nC if nColumns is None else int(nColumns)
        _g = globals()["_f"]

        def _f(shape, dtype=np.float64):
            return _g(shape, alloc, dtype)
        self.DMS_tracers = _f((nL, nC, abi.DMS_TRACER_CNT))
        self.cell_thickness = _f((nL, nC))
        self.number_of_active_levels = _f((nC,), np.int32)
        self.forcing = {n: _f((nC,)) for n in abi.DMS_FORCING_C1}
        self.forcing["netFlux"] = _f((nC, abi.DMS_TRACER_CNT))
        self.lcalc_DMS_gas_flux = 1
        self.DMS_tendencies = _f((nL, nC, abi.DMS_TRACER_CNT))
        self.diag = {n: _f((nL, nC)) for n in abi.DMS_DIAG} if diagnostics else {}
        self.flux_diag = {n: _f((nC,)) for n in abi.DMS_FLUX_DIAG}

    def c_input(self):
        s = abi.DmsInput()
        s.DMS_tracers = abi.fptr(self.DMS_tracers)
        s.cell_thickness = abi.fptr(self.cell_thickness)
        s.number_of_active_levels = abi.iptr(self.number_of_active_levels)
        return s

    def c_forcing(self):
        s = abi.DmsForcing()
        for n, a in self.forcing.items():
            setattr(s, n, abi.fptr(a))
        s.lcalc_DMS_gas_flux = int(self.lcalc_DMS_gas_flux)
        return s

    def c_output(self):
        s = abi.DmsOutput()
        s.DMS_tendencies = abi.fptr(self.DMS_tendencies)
        return s

    def c_diag(self, enabled=True):
        s = abi.DmsDiagnostics()
        if enabled:
            for n, a in self.diag.items():
                setattr(s, n, abi.fptr(a))
        return s

    def c_flux_diag(self):
        s = abi.DmsFluxDiagnostics()
        for n, a in self.flux_diag.items():
            setattr(s, n, abi.fptr(a))
        return s

    def copy(self):
        o = DmsColumns(self.nLevelsMax, self.nColumnsMax, self.nColumns, diagnostics=bool(self.diag))
        for n in ("DMS_tracers", "cell_thickness", "number_of_active_levels", "DMS_tendencies"):
            getattr(o, n)[...] = getattr(self, n)
        for d in ("forcing", "diag", "flux_diag"):
            for n, a in getattr(self, d).items():
                getattr(o, d)[n][...] = a
        o.lcalc_DMS_gas_flux = self.lcalc_DMS_gas_flux
        return o


class MacrosColumns:
    def __init__(self, nLevelsMax, nColumnsMax, nColumns=None, alloc=None, diagnostics=True):
        nL, nC = int(nLevelsMax), int(nColumnsMax)
        self.nLevelsMax, self.nColumnsMax = nL, nC
        self.nColumns = nC if nColumns is None else int(nColumns)
        _g = globals()["_f"]

        def _f(shape, dtype=np.float64):
            return _g(shape, alloc, dtype)
        self.MACROS_tracers = _f((nL, nC, abi.MACROS_TRACER_CNT))
        self.cell_thickness = _f((nL, nC))
        self.number_of_active_levels = _f((nC,), np.int32)
        self.MACROS_tendencies = _f((nL, nC, abi.MACROS_TRACER_CNT))
        self.diag = {n: _f((nL, nC)) for n in abi.MACROS_DIAG} if diagnostics else {}

    def c_input(self):
        s = abi.MacrosInput()
        s.MACROS_tracers = abi.fptr(self.MACROS_tracers)
        s.cell_thickness = abi.fptr(self.cell_thickness)
        s.number_of_active_levels = abi.iptr(self.number_of_active_levels)
        return s

    def c_output(self):
        s = abi.MacrosOutput()
        s.MACROS_tendencies = abi.fptr(self.MACROS_tendencies)
        return s

    def c_diag(self, enabled=True):
        s = abi.MacrosDiagnostics()
        if enabled:
            for n, a in self.diag.items():
                setattr(s, n, abi.fptr(a))
        return s

    def copy(self):
        o = MacrosColumns(self.nLevelsMax, self.nColumnsMax, self.nColumns, diagnostics=bool(self.diag))
        for n in ("MACROS_tracers", "cell_thickness", "number_of_active_levels",
                  "MACROS_tendencies"):
            getattr(o, n)[...] = getattr(self, n)
        for n, a in self.diag.items():
            o.diag[n][...] = a
        return o


# ------------------------------------------------------------------ synthetic data

class _SynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("nLevelsMax", C.c_int), ("nColumnsMax", C.c_int),
                ("nColumns", C.c_int), ("column0", C.c_longlong), ("nlev_active", C.c_int),
                ("ragged", C.c_int), ("soa", C.c_int), ("jitter", C.c_int),
                ("nthreads", C.c_int)]


_synth_lib = None


def synth_lib():
    global _synth_lib
    if _synth_lib is None:
        import os
        path = os.path.join(abi.HERE, "csrc", "libbgc_synth.so")
        if not os.path.exists(path):
            raise RuntimeError("%s missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
        _synth_lib = C.CDLL(path)
        _synth_lib.bgc_synth_fill.restype = C.c_int
        _synth_lib.bgc_synth_co2_points.restype = C.c_int
    return _synth_lib


SEED_COLUMNS = 0x0B6C0003
SEED_CO2 = 0x0B6C0001


def synth_fill(bgc=None, dms=None, macros=None, *, bgc_ind, dms_ind=None, macros_ind=None,
               seed=SEED_COLUMNS, column0=0, nlev_active=None, ragged=False, jitter=True,
               nthreads=0, soa=False):
    """Fill the *input* and *forcing* members of the given containers in place
    (Fortran layout) with the synthetic columns of SURVEY.md 8(d)."""
    ref = bgc or dms or macros
    sp = _SynthSpec(seed, ref.nLevelsMax, ref.nColumnsMax, ref.nColumns, column0,
                    nlev_active or ref.nLevelsMax, int(ragged), int(soa), int(jitter), nthreads)
    keep = []

    def byref_or_null(s):
        if s is None:
            return None
        keep.append(s)
        return C.byref(s)

    args = [C.byref(sp), C.byref(bgc_ind), byref_or_null(dms_ind), byref_or_null(macros_ind),
            byref_or_null(bgc.c_input() if bgc else None),
            byref_or_null(bgc.c_forcing() if bgc else None),
            byref_or_null(dms.c_input() if dms else None),
            byref_or_null(dms.c_forcing() if dms else None),
            byref_or_null(macros.c_input() if macros else None)]
    rc = synth_lib().bgc_synth_fill(*args)
    if rc != 0:
        raise RuntimeError("bgc_synth_fill failed: %d" % rc)


def synth_co2_points(n, seed=SEED_CO2, i0=0):
    """Config 2 inputs: dict of 11 arrays of length n for co2calc_1point."""
    out = np.zeros((11, n), dtype=np.float64)
    rc = synth_lib().bgc_synth_co2_points(C.c_uint64(seed), C.c_longlong(i0), C.c_int(n),
                                          abi.dptr(out))
    if rc != 0:
        raise RuntimeError("bgc_synth_co2_points failed: %d" % rc)
    names = ["depth", "temp", "salt", "dic", "ta", "pt", "sit", "phlo", "phhi", "xco2", "atmpres"]
    return {k: out[i].copy() for i, k in enumerate(names)}


def synth_fill_device(parms, bgc, dms, mac, column0=0, *, ragged=False, jitter=True, nthreads=0,
                      seed=SEED_COLUMNS):
    """Generate the synthetic columns [column0, column0 + nColumnsMax) directly in the SoA layout on
    the host (no Fortran-layout intermediate) and copy the INPUT members into the device
    containers (host.DeviceBgcColumns / DeviceDmsColumns / DeviceMacrosColumns).  The generator
    is keyed by the global column index, so any sharding sees the same columns.  Returns the
    number of active cells."""
    import torch
    nL, nC = bgc.nLevelsMax, bgc.nColumnsMax
    h = {}

    def arr(shape, dtype=np.float64):
        return np.zeros(shape, dtype=dtype)
    h["tr"] = arr((abi.BGC_TRACER_CNT, nL, nC))
    for n in bgc.K2_IN:
        h[n] = arr((nL, nC))
    h["lat"] = arr((nC,)); h["kmax"] = arr((nC,), np.int32)
    hf = {n: arr((nL, nC)) for n in ("FESEDFLUX",)}
    hf.update({n: arr((nC,)) for n in abi.BGC_FORCING_C1})
    hf.update({n: arr((abi.BGC_TRACER_CNT, nC)) for n in abi.BGC_FORCING_FLUX})
    h["dtr"] = arr((abi.DMS_TRACER_CNT, nL, nC)); h["ddz"] = arr((nL, nC)); h["dkmax"] = arr((nC,), np.int32)
    hdf = {n: arr((nC,)) for n in abi.DMS_FORCING_C1}
    hdf["netFlux"] = arr((abi.DMS_TRACER_CNT, nC))
    h["mtr"] = arr((abi.MACROS_TRACER_CNT, nL, nC)); h["mdz"] = arr((nL, nC)); h["mkmax"] = arr((nC,), np.int32)

    cin = abi.BgcInput()
    cin.BGC_tracers = abi.dptr(h["tr"])
    for n in bgc.K2_IN:
        setattr(cin, n, abi.dptr(h[n]))
    cin.cell_latitude = abi.dptr(h["lat"]); cin.number_of_active_levels = abi.iptr(h["kmax"])
    cfo = abi.BgcForcing()
    for n, a in hf.items():
        setattr(cfo, n, abi.dptr(a))
    din = abi.DmsInput()
    din.DMS_tracers = abi.dptr(h["dtr"]); din.cell_thickness = abi.dptr(h["ddz"])
    din.number_of_active_levels = abi.iptr(h["dkmax"])
    dfo = abi.DmsForcing()
    for n, a in hdf.items():
        setattr(dfo, n, abi.dptr(a))
    min_ = abi.MacrosInput()
    min_.MACROS_tracers = abi.dptr(h["mtr"]); min_.cell_thickness = abi.dptr(h["mdz"])
    min_.number_of_active_levels = abi.iptr(h["mkmax"])

    sp = _SynthSpec(seed, nL, nC, bgc.nColumns, column0, nL, int(ragged), 1, int(jitter), nthreads)
    rc = synth_lib().bgc_synth_fill(C.byref(sp), C.byref(parms.ind), C.byref(parms.dms_ind),
                                    C.byref(parms.macros_ind), C.byref(cin), C.byref(cfo), C.byref(din),
                                    C.byref(dfo), C.byref(min_))
    if rc != 0:
        raise RuntimeError("bgc_synth_fill failed: %d" % rc)

    def put(t, a):
        t.copy_(torch.from_numpy(a))
    put(bgc.BGC_tracers, h["tr"])
    for n in bgc.K2_IN:
        put(getattr(bgc, n), h[n])
    put(bgc.cell_latitude, h["lat"]); put(bgc.number_of_active_levels, h["kmax"])
    for n, a in hf.items():
        put(bgc.forcing[n], a)
    if dms is not None:
        put(dms.DMS_tracers, h["dtr"]); put(dms.cell_thickness, h["ddz"]); put(dms.number_of_active_levels, h["dkmax"])
        for n, a in hdf.items():
            put(dms.forcing[n], a)
    if mac is not None:
        put(mac.MACROS_tracers, h["mtr"]); put(mac.cell_thickness, h["mdz"]); put(mac.number_of_active_levels, h["mkmax"])
    if bgc.BGC_tracers.is_cuda:
        torch.cuda.synchronize()
    kmax = h["kmax"].astype(np.int64)
    kmax[bgc.nColumns:] = 0
    return int(kmax.sum())
