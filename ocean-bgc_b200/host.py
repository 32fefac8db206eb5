"""Host-side mirror of the reference's public procedures over the C ABI.

Same names, argument meaning and side effects as the Fortran entry points:

    BGC_parms_init / BGC_init        BGC_parms.F90:497, BGC_mod.F90:184
    BGC_SourceSink                   BGC_mod.F90:340
    BGC_SurfaceFluxes                BGC_mod.F90:2706
    co2calc_1point (batched)         co2calc.F90:75
    DMS_SourceSink / DMS_SurfaceFluxes   DMS_mod.F90:156, :778
    MACROS_SourceSink                MACROS_mod.F90:137

Everything here calls into ocean-bgc_b200/csrc/libbgc_b200.so (CUDA, sm_100a).
There is NO CPU fallback: a missing library or a missing GPU raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .columns import BgcColumns, DmsColumns, MacrosColumns

CSRC = os.path.join(abi.HERE, "csrc")
LIB_NAME = {"prod": "libbgc_b200.so", "strict": "libbgc_b200_strict.so"}

# every symbol include/bgc_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "bgc_last_error", "bgc_version", "bgc_parms_init", "bgc_default_tracer_indices", "bgc_init",
    "dms_parms_init", "dms_default_tracer_indices", "macros_parms_init",
    "macros_default_tracer_indices", "bgc_ctx_create", "bgc_ctx_destroy", "bgc_ctx_set_stream",
    "bgc_ctx_synchronize", "bgc_get_status", "bgc_set_params", "dms_set_params",
    "macros_set_params", "bgc_source_sink", "bgc_surface_fluxes", "bgc_co2calc_points",
    "dms_source_sink", "dms_surface_fluxes", "macros_source_sink", "bgc_inventory_enable",
    "bgc_inventory_reset", "bgc_inventory_get", "bgc_inventory_device_ptr", "bgc_comm_unique_id",
    "bgc_comm_init_rank", "bgc_inventory_allreduce", "bgc_host_alloc", "bgc_host_free",
    "bgc_host_register", "bgc_host_unregister", "bgc_layout_to_soa", "bgc_layout_to_fortran",
    "bgc_timing_enable", "bgc_timing_reset", "bgc_timing_get", "bgc_kernel_name",
    "bgc_ctx_set_deferred_join", "bgc_carbonate_join", "bgc_ctx_set_concurrency",
    "bgc_diag_accumulate_enable", "bgc_diag_flush", "bgc_layout_mpas_to_soa", "bgc_layout_soa_to_mpas",
    "bgc_inventory_allreduce_begin", "bgc_inventory_allreduce_end",
    "bgc_graph_capture_begin", "bgc_graph_capture_end", "bgc_graph_launch", "bgc_graph_destroy",
    "bgc_ctx_set_zero_shortcut", "bgc_comp_co3terms", "bgc_comp_co3_sat_vals",
    "bgc_layout_soa_to_mpas_weighted", "bgc_state_device_ptr", "bgc_state_set", "bgc_state_get",
    "bgc_transfer_bytes",
]


class BgcError(RuntimeError):
    pass


_libs = {}


def lib(flavour=None):
    """The C-ABI library.  flavour: "prod" (FMA contraction on) or "strict"
    (-fmad=false parity build); default from $BGC_B200_FLAVOUR, else "prod"."""
    flavour = flavour or os.environ.get("BGC_B200_FLAVOUR", "prod")
    if flavour not in _libs:
        # $BGC_B200_LIBDIR: another build of the libraries (A/B runs against an older tree; symbols it lacks are skipped)
        libdir = os.environ.get("BGC_B200_LIBDIR")
        path = os.path.join(libdir or CSRC, LIB_NAME[flavour])
        if not os.path.exists(path):
            raise BgcError("%s is missing: build it with `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (there is no CPU fallback)" % path)
        L = C.CDLL(path)
        L.bgc_last_error.restype = C.c_char_p
        L.bgc_version.restype = C.c_char_p
        L.bgc_kernel_name.restype = C.c_char_p
        for s in ABI_SYMBOLS:
            if s not in ("bgc_last_error", "bgc_version", "bgc_kernel_name"):
                if libdir and not hasattr(L, s):
                    continue
                getattr(L, s).restype = C.c_int
        _libs[flavour] = L
    return _libs[flavour]


def check(L, rc):
    if rc != abi.BGC_OK:
        raise BgcError("C ABI call failed (%d): %s" % (rc, L.bgc_last_error().decode()))


class Parms:
    """The parameter tables as *_parms_init produce them, plus host-chosen tracer
    slots (declaration order unless `permute_tracers` is used)."""

    def __init__(self, flavour=None):
        L = lib(flavour)
        self.bgc = abi.BgcParams()
        self.autotrophs = abi.BgcAutotroph4()
        self.ind = abi.BgcIndices()
        self.dms = abi.DmsParams()
        self.dms_ind = abi.DmsIndices()
        self.macros = abi.MacrosParams()
        self.macros_ind = abi.MacrosIndices()
        check(L, L.bgc_parms_init(C.byref(self.bgc), self.autotrophs, C.byref(self.ind)))
        check(L, L.bgc_default_tracer_indices(C.byref(self.ind)))
        check(L, L.bgc_init(C.byref(self.ind), self.autotrophs))
        check(L, L.dms_parms_init(C.byref(self.dms)))
        check(L, L.dms_default_tracer_indices(C.byref(self.dms_ind)))
        check(L, L.macros_parms_init(C.byref(self.macros)))
        check(L, L.macros_default_tracer_indices(C.byref(self.macros_ind)))
        self._L = L

    def permute_tracers(self, perm):
        for i, (n, _) in enumerate(abi.BgcIndices._fields_[:abi.BGC_TRACER_CNT]):
            setattr(self.ind, n, int(perm[i]) + 1)
        check(self._L, self._L.bgc_init(C.byref(self.ind), self.autotrophs))


BGC_parms_init = Parms   # reference-named alias


class Context:
    """One GPU + its persistent device arena (bgc_ctx)."""

    def __init__(self, nLevelsMax, nColumnsMax, device=0, flavour=None, parms=None):
        self.L = lib(flavour)
        self.ptr = C.c_void_p()
        check(self.L, self.L.bgc_ctx_create(C.c_int(device), C.c_int(nLevelsMax), C.c_int(nColumnsMax),
                                            C.byref(self.ptr)))
        self.device = device
        self.nLevelsMax, self.nColumnsMax = nLevelsMax, nColumnsMax
        if parms is not None:
            self.set_params(parms)

    def set_params(self, parms):
        L = self.L
        check(L, L.bgc_set_params(self.ptr, C.byref(parms.bgc), parms.autotrophs, C.byref(parms.ind)))
        check(L, L.dms_set_params(self.ptr, C.byref(parms.dms), C.byref(parms.dms_ind)))
        check(L, L.macros_set_params(self.ptr, C.byref(parms.macros), C.byref(parms.macros_ind)))
        self.parms = parms

    def set_stream(self, cuda_stream):
        check(self.L, self.L.bgc_ctx_set_stream(self.ptr, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        check(self.L, self.L.bgc_ctx_synchronize(self.ptr))

    def set_deferred_join(self, on=True):
        """Defer the join of the carbonate side stream to the next join point (bgc_b200.h)."""
        check(self.L, self.L.bgc_ctx_set_deferred_join(self.ptr, C.c_int(int(on))))

    def set_concurrency(self, on=True):
        check(self.L, self.L.bgc_ctx_set_concurrency(self.ptr, C.c_int(int(on))))

    def diag_accumulate(self, on=True):
        """Host-layout calls add their diagnostics into device accumulators instead of downloading
        them (bgc_b200.h, "Diagnostics accumulation")."""
        check(self.L, self.L.bgc_diag_accumulate_enable(self.ptr, C.c_int(int(on))))

    def diag_flush(self, bgc=None, dms=None, macros=None, scale=1.0, reset=True):
        """Download scale * accumulated sums into the diag arrays of the given host containers."""
        ref = bgc or dms or macros
        nL, nC = ref.nLevelsMax, ref.nColumnsMax
        keep = [x.c_diag(True) if x is not None else None for x in (bgc, dms, macros)]
        args = [C.byref(k) if k is not None else None for k in keep]
        check(self.L, self.L.bgc_diag_flush(self.ptr, args[0], args[1], args[2], C.c_int(nL), C.c_int(nC),
                                            C.c_double(scale), C.c_int(int(reset))))

    def mpas_to_soa(self, dev_mpas, dev_soa, slot_of_tracer, nL, nC):
        """T(tracer,k,cell) device array -> SoA device array (raw device addresses)."""
        m = (C.c_int * len(slot_of_tracer))(*[int(x) for x in slot_of_tracer])
        check(self.L, self.L.bgc_layout_mpas_to_soa(self.ptr, abi.raw_dptr(dev_mpas), abi.raw_dptr(dev_soa),
                                                    C.c_int(len(slot_of_tracer)), m, C.c_int(nL), C.c_int(nC)))

    def soa_to_mpas(self, dev_soa, dev_mpas, slot_of_tracer, nL, nC, alpha=1.0, beta=0.0, dev_weight=None):
        """T(n,k,cell) = beta*T + alpha*w(k,cell)*soa(cell,k,slot[n]): layout change fused with the tracer
        update; dev_weight = device address of w(k,cell) (MPAS layerThickness layout) or None for w = 1."""
        m = (C.c_int * len(slot_of_tracer))(*[int(x) for x in slot_of_tracer])
        check(self.L, self.L.bgc_layout_soa_to_mpas_weighted(
            self.ptr, abi.raw_dptr(dev_soa), abi.raw_dptr(dev_mpas), C.c_int(len(slot_of_tracer)), m, C.c_int(nL),
            C.c_int(nC), C.c_double(alpha), C.c_double(beta), abi.raw_dptr(dev_weight or 0)))

    # ---- device-resident model state (restart fields): bgc_b200.h "Device-resident model state"
    STATE = {"PH_PREV_3D": 0, "PH_PREV_ALT_CO2_3D": 1, "surface_pH": 2, "surface_pH_alt_co2": 3}

    def state_device_ptr(self, name, nL, nC):
        p = C.POINTER(C.c_double)()
        check(self.L, self.L.bgc_state_device_ptr(self.ptr, C.c_int(self.STATE[name]), C.c_int(nL), C.c_int(nC),
                                                  C.byref(p)))
        return C.cast(p, C.c_void_p).value

    def state_set(self, name, host_array, nL, nC):
        a = np.asfortranarray(host_array, dtype=np.float64)
        check(self.L, self.L.bgc_state_set(self.ptr, C.c_int(self.STATE[name]), abi.fptr(a), C.c_int(nL), C.c_int(nC)))

    def state_get(self, name, nL, nC):
        shape = (nL, nC) if self.STATE[name] < 2 else (nC,)
        a = np.zeros(shape, dtype=np.float64, order="F")
        check(self.L, self.L.bgc_state_get(self.ptr, C.c_int(self.STATE[name]), abi.fptr(a), C.c_int(nL), C.c_int(nC)))
        return a

    def set_zero_shortcut(self, on=True):
        check(self.L, self.L.bgc_ctx_set_zero_shortcut(self.ptr, C.c_int(int(on))))

    def transfer_bytes(self, reset=False):
        """(host -> device, device -> host) bytes the host-layout calls of this ctx have copied"""
        b = (C.c_ulonglong * 2)()
        check(self.L, self.L.bgc_transfer_bytes(self.ptr, b, C.c_int(int(reset))))
        return int(b[0]), int(b[1])

    def carbonate_join(self):
        check(self.L, self.L.bgc_carbonate_join(self.ptr))

    def status(self, reset=False):
        st = abi.BgcStatus()
        check(self.L, self.L.bgc_get_status(self.ptr, C.byref(st), C.c_int(int(reset))))
        return abi.struct_to_dict(st)

    def timing_enable(self, on=True):
        check(self.L, self.L.bgc_timing_enable(self.ptr, C.c_int(int(on))))

    def timing_reset(self):
        check(self.L, self.L.bgc_timing_reset(self.ptr))

    def timing(self):
        """{kernel name: (device ms under timing, timed launches, all launches)}"""
        out = {}
        for kid in range(abi.DEFINES["BGC_KERNEL_ID_COUNT"]):
            ms, tl, nl = C.c_double(), C.c_ulonglong(), C.c_ulonglong()
            check(self.L, self.L.bgc_timing_get(self.ptr, C.c_int(kid), C.byref(ms), C.byref(tl), C.byref(nl)))
            out[self.L.bgc_kernel_name(C.c_int(kid)).decode()] = (ms.value, tl.value, nl.value)
        return out

    def launch_count(self):
        return sum(v[2] for v in self.timing().values())

    def inventory_enable(self, on=True):
        check(self.L, self.L.bgc_inventory_enable(self.ptr, C.c_int(int(on))))

    def inventory_reset(self):
        check(self.L, self.L.bgc_inventory_reset(self.ptr))

    def inventory_get(self):
        out = (C.c_double * abi.BGC_INVENTORY_LEN)()
        check(self.L, self.L.bgc_inventory_get(self.ptr, out))
        return np.array(out[:])

    def inventory_allreduce(self):
        out = (C.c_double * abi.BGC_INVENTORY_LEN)()
        check(self.L, self.L.bgc_inventory_allreduce(self.ptr, out))
        return np.array(out[:])

    def inventory_allreduce_begin(self):
        check(self.L, self.L.bgc_inventory_allreduce_begin(self.ptr))

    def inventory_allreduce_end(self):
        out = (C.c_double * abi.BGC_INVENTORY_LEN)()
        check(self.L, self.L.bgc_inventory_allreduce_end(self.ptr, out))
        return np.array(out[:], dtype=np.float64)

    def graph_capture_begin(self):
        check(self.L, self.L.bgc_graph_capture_begin(self.ptr))

    def graph_capture_end(self):
        g = C.c_void_p()
        check(self.L, self.L.bgc_graph_capture_end(self.ptr, C.byref(g)))
        return g

    def graph_launch(self, g):
        check(self.L, self.L.bgc_graph_launch(self.ptr, g))

    def graph_destroy(self, g):
        check(self.L, self.L.bgc_graph_destroy(g))

    def comm_unique_id(self):
        buf = (C.c_ubyte * 128)()
        check(self.L, self.L.bgc_comm_unique_id(buf))
        return bytes(buf)

    def comm_init_rank(self, nranks, rank, uid):
        buf = (C.c_ubyte * 128).from_buffer_copy(uid)
        check(self.L, self.L.bgc_comm_init_rank(self.ptr, C.c_int(nranks), C.c_int(rank), buf))

    def close(self):
        if self.ptr:
            self.L.bgc_ctx_destroy(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _dims(cols):
    return C.c_int(cols.nLevelsMax), C.c_int(cols.nColumnsMax), C.c_int(cols.nColumns)


def _space(cols):
    return C.c_int(getattr(cols, "mem_space", abi.BGC_MEM_HOST_FORTRAN))


def BGC_SourceSink(ctx, cols, alt_co2_use_eco=True, diagnostics=True):
    """BGC_SourceSink(autotrophs, BGC_indices, BGC_input, BGC_forcing, BGC_output,
    BGC_diagnostic_fields, numLevelsMax, numColumnsMax, numColumns, alt_co2_use_eco)
    — the tables come from ctx.set_params; `cols` carries the four derived types."""
    cin, cfo, cout, cdg = cols.c_input(), cols.c_forcing(), cols.c_output(), cols.c_diag(diagnostics)
    nL, nC, n = _dims(cols)
    check(ctx.L, ctx.L.bgc_source_sink(ctx.ptr, C.byref(cin), C.byref(cfo), C.byref(cout), C.byref(cdg),
                                       nL, nC, n, C.c_int(int(alt_co2_use_eco)), _space(cols)))


def BGC_SurfaceFluxes(ctx, cols):
    cin, cfo, cfd = cols.c_input(), cols.c_forcing(), cols.c_flux_diag()
    nL, nC, n = _dims(cols)
    check(ctx.L, ctx.L.bgc_surface_fluxes(ctx.ptr, C.byref(cin), C.byref(cfo), C.byref(cfd), nL, nC, n,
                                          _space(cols)))


def DMS_SourceSink(ctx, cols, diagnostics=True):
    cin, cfo, cout, cdg = cols.c_input(), cols.c_forcing(), cols.c_output(), cols.c_diag(diagnostics)
    nL, nC, n = _dims(cols)
    check(ctx.L, ctx.L.dms_source_sink(ctx.ptr, C.byref(cin), C.byref(cfo), C.byref(cout), C.byref(cdg),
                                       nL, nC, n, _space(cols)))


def DMS_SurfaceFluxes(ctx, cols):
    cin, cfo, cfd = cols.c_input(), cols.c_forcing(), cols.c_flux_diag()
    nL, nC, n = _dims(cols)
    check(ctx.L, ctx.L.dms_surface_fluxes(ctx.ptr, C.byref(cin), C.byref(cfo), C.byref(cfd), nL, nC, n,
                                          _space(cols)))


def MACROS_SourceSink(ctx, cols, diagnostics=True):
    cin, cout, cdg = cols.c_input(), cols.c_output(), cols.c_diag(diagnostics)
    nL, nC, n = _dims(cols)
    check(ctx.L, ctx.L.macros_source_sink(ctx.ptr, C.byref(cin), C.byref(cout), C.byref(cdg), nL, nC, n,
                                          _space(cols)))


_PT_IN = ("depth", "temp", "salt", "dic", "ta", "pt", "sit", "phlo", "phhi", "xco2", "atmpres")
_PT_OUT = ("ph", "co2star", "dco2star", "pco2surf", "dpco2")


def co2calc_points(ctx, pts):
    """Batched co2calc_1point(depth, .true., .true., temp, salt, dic, ta, pt, sit, phlo, phhi,
    ph, xco2, atmpres, co2star, dco2star, pCO2surf, dpco2) over host numpy arrays."""
    n = len(pts["temp"])
    a = {k: np.ascontiguousarray(pts[k], dtype=np.float64) for k in _PT_IN}
    out = {k: np.zeros(n) for k in _PT_OUT}
    check(ctx.L, ctx.L.bgc_co2calc_points(ctx.ptr, C.c_int(n), *[abi.dptr(a[k]) for k in _PT_IN],
                                          *[abi.dptr(out[k]) for k in _PT_OUT],
                                          C.c_int(abi.BGC_MEM_HOST_FORTRAN)))
    return out


def comp_CO3terms_points(ctx, k, depth, temp, salt, dic, ta, pt, sit, phlo, phhi):
    """Batched comp_CO3terms(k, depth, .true., temp, salt, dic_in, ta_in, pt_in, sit_in, phlo, phhi, ph,
    H2CO3, HCO3, CO3) (co2calc.F90:214) over host numpy arrays; k = array of 1-based level indices or an int."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (depth, temp, salt, dic, ta, pt, sit, phlo, phhi)]
    n = len(arrs[1])
    out = {nm: np.zeros(n) for nm in ("pH", "H2CO3", "HCO3", "CO3")}
    if np.ndim(k) == 0:
        kp, k_all, keep = C.POINTER(C.c_int)(), int(k), None
    else:
        keep = np.ascontiguousarray(k, dtype=np.int32)
        kp, k_all = abi.iptr(keep), 1
    check(ctx.L, ctx.L.bgc_comp_co3terms(ctx.ptr, C.c_int(n), kp, C.c_int(k_all), *[abi.dptr(a) for a in arrs],
                                         *[abi.dptr(out[nm]) for nm in ("pH", "H2CO3", "HCO3", "CO3")],
                                         C.c_int(abi.BGC_MEM_HOST_FORTRAN)))
    return out


def comp_co3_sat_vals_points(ctx, k, depth, temp, salt):
    """Batched comp_co3_sat_vals(k, depth, temp, salt, co3_sat_calc, co3_sat_arag) (co2calc.F90:1096)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (depth, temp, salt)]
    n = len(arrs[1])
    calc, arag = np.zeros(n), np.zeros(n)
    if np.ndim(k) == 0:
        kp, k_all, keep = C.POINTER(C.c_int)(), int(k), None
    else:
        keep = np.ascontiguousarray(k, dtype=np.int32)
        kp, k_all = abi.iptr(keep), 1
    check(ctx.L, ctx.L.bgc_comp_co3_sat_vals(ctx.ptr, C.c_int(n), kp, C.c_int(k_all), *[abi.dptr(a) for a in arrs],
                                             abi.dptr(calc), abi.dptr(arag), C.c_int(abi.BGC_MEM_HOST_FORTRAN)))
    return calc, arag


def co2calc_points_device(ctx, dev_in, dev_out, n):
    """Same on device arrays: dev_in / dev_out are dicts of raw device addresses."""
    check(ctx.L, ctx.L.bgc_co2calc_points(ctx.ptr, C.c_int(n), *[abi.raw_dptr(dev_in[k]) for k in _PT_IN],
                                          *[abi.raw_dptr(dev_out[k]) for k in _PT_OUT],
                                          C.c_int(abi.BGC_MEM_DEVICE_SOA)))


# ------------------------------------------------------------------ device-resident containers
# torch is used for device memory only (allocation, host<->device staging of
# test data); every computation goes through the C ABI above.

def _torch():
    import torch
    return torch


def _sync(device):
    """(containers on device "cpu" exist only for dry runs of the test harness)"""
    torch = _torch()
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


class _DeviceMixin:
    mem_space = abi.BGC_MEM_DEVICE_SOA

    def _alloc(self, shape, dtype=None):
        torch = _torch()
        return torch.zeros(shape, dtype=dtype or torch.float64, device=self.device)

    @staticmethod
    def _soa(a):
        """numpy Fortran-layout (k,col[,n]) -> contiguous (n,k,col) / (k,col)."""
        if a.ndim == 3:
            return np.ascontiguousarray(np.transpose(a, (2, 0, 1)))
        if a.ndim == 2:
            return np.ascontiguousarray(a)
        return np.ascontiguousarray(a)

    @staticmethod
    def _flux(a):
        """(col,n) Fortran -> contiguous (n,col)."""
        return np.ascontiguousarray(a.T) if a.ndim == 2 else np.ascontiguousarray(a)


class DeviceBgcColumns(_DeviceMixin):
    """The BGC derived types resident on the GPU in the column-fastest SoA layout
    (A(k,col,n) at col + nC*(k + nL*n))."""

    K2_IN = ("PotentialTemperature", "Salinity", "cell_center_depth", "cell_thickness", "cell_bottom_depth")

    def __init__(self, nLevelsMax, nColumnsMax, nColumns=None, device="cuda:0", diagnostics=True):
        torch = _torch()
        self.device = torch.device(device)
        nL, nC = int(nLevelsMax), int(nColumnsMax)
        self.nLevelsMax, self.nColumnsMax = nL, nC
        self.nColumns = nC if nColumns is None else int(nColumns)
        A = self._alloc
        self.BGC_tracers = A((abi.BGC_TRACER_CNT, nL, nC))
        for n in self.K2_IN:
            setattr(self, n, A((nL, nC)))
        self.cell_latitude = A((nC,))
        self.number_of_active_levels = A((nC,), torch.int32)
        self.forcing = {}
        for n in abi.BGC_FORCING_K2[:1]:   # FESEDFLUX; restoring fields only when lrest_* is used
            self.forcing[n] = A((nL, nC))
        for n in abi.BGC_FORCING_C1:
            self.forcing[n] = A((nC,))
        for n in abi.BGC_FORCING_FLUX:
            self.forcing[n] = A((abi.BGC_TRACER_CNT, nC))
        self.lcalc_O2_gas_flux = 1
        self.lcalc_CO2_gas_flux = 1
        self.BGC_tendencies = A((abi.BGC_TRACER_CNT, nL, nC))
        self.PH_PREV_3D = A((nL, nC))
        self.PH_PREV_ALT_CO2_3D = A((nL, nC))
        self.diag = {}
        if diagnostics:
            for n in abi.BGC_DIAG_K2:
                self.diag[n] = A((nL, nC))
            for n in abi.BGC_DIAG_KA:
                self.diag[n] = A((abi.BGC_AUTOTROPH_CNT, nL, nC))
            for n in abi.BGC_DIAG_CA:
                self.diag[n] = A((abi.BGC_AUTOTROPH_CNT, nC))
            for n in abi.BGC_DIAG_C1:
                self.diag[n] = A((nC,))
        self.flux_diag = {n: A((nC,)) for n in abi.BGC_FLUX_DIAG}

    def nbytes(self):
        t = [self.BGC_tracers, self.BGC_tendencies, self.PH_PREV_3D, self.PH_PREV_ALT_CO2_3D,
             self.cell_latitude, self.number_of_active_levels] + [getattr(self, n) for n in self.K2_IN]
        t += list(self.forcing.values()) + list(self.diag.values()) + list(self.flux_diag.values())
        return sum(x.numel() * x.element_size() for x in t)

    def load(self, host):
        """Copy a host BgcColumns (Fortran layout) into this container."""
        torch = _torch()

        def put(dst, arr):
            dst.copy_(torch.from_numpy(arr))
        put(self.BGC_tracers, self._soa(host.BGC_tracers))
        for n in self.K2_IN:
            put(getattr(self, n), self._soa(getattr(host, n)))
        put(self.cell_latitude, host.cell_latitude)
        put(self.number_of_active_levels, host.number_of_active_levels)
        for n in abi.BGC_FORCING_K2[1:]:   # nutrient-restoring fields: resident only when the host uses them
            if n not in self.forcing and np.any(host.forcing[n]):
                self.forcing[n] = self._alloc((self.nLevelsMax, self.nColumnsMax))
        for n, t in self.forcing.items():
            a = host.forcing[n]
            put(t, self._flux(a) if n in abi.BGC_FORCING_FLUX else self._soa(a))
        put(self.BGC_tendencies, self._soa(host.BGC_tendencies))
        put(self.PH_PREV_3D, self._soa(host.PH_PREV_3D))
        put(self.PH_PREV_ALT_CO2_3D, self._soa(host.PH_PREV_ALT_CO2_3D))
        for n, t in self.diag.items():
            a = host.diag[n]
            put(t, self._flux(a) if n in abi.BGC_DIAG_CA else self._soa(a))
        for n, t in self.flux_diag.items():
            put(t, host.flux_diag[n])
        self.nColumns = host.nColumns
        self.lcalc_O2_gas_flux, self.lcalc_CO2_gas_flux = host.lcalc_O2_gas_flux, host.lcalc_CO2_gas_flux
        _sync(self.device)   # torch copies run on torch's stream, library calls on the ctx stream
        return self

    def store(self, host):
        """Copy outputs / in-out members back into a host BgcColumns."""
        def get(t, like, per_column_n=False):
            a = t.cpu().numpy()
            if like.ndim == 3:
                like[...] = np.transpose(a, (1, 2, 0))     # (n,k,col) -> (k,col,n)
            elif per_column_n:
                like[...] = a.T                            # (n,col) -> (col,n), whatever the extents are
            else:
                like[...] = a
        get(self.BGC_tendencies, host.BGC_tendencies)
        get(self.PH_PREV_3D, host.PH_PREV_3D)
        get(self.PH_PREV_ALT_CO2_3D, host.PH_PREV_ALT_CO2_3D)
        for n, t in self.diag.items():
            get(t, host.diag[n], n in abi.BGC_DIAG_CA)
        for n, t in self.flux_diag.items():
            get(t, host.flux_diag[n])
        for n, t in self.forcing.items():
            get(t, host.forcing[n], n in abi.BGC_FORCING_FLUX)
        return host

    # ---- C-ABI argument blocks
    def c_input(self):
        s = abi.BgcInput()
        s.BGC_tracers = abi.raw_dptr(self.BGC_tracers.data_ptr())
        for n in self.K2_IN:
            setattr(s, n, abi.raw_dptr(getattr(self, n).data_ptr()))
        s.cell_latitude = abi.raw_dptr(self.cell_latitude.data_ptr())
        s.number_of_active_levels = abi.raw_iptr(self.number_of_active_levels.data_ptr())
        return s

    def c_forcing(self):
        s = abi.BgcForcing()
        for n, t in self.forcing.items():
            setattr(s, n, abi.raw_dptr(t.data_ptr()))
        s.lcalc_O2_gas_flux = int(self.lcalc_O2_gas_flux)
        s.lcalc_CO2_gas_flux = int(self.lcalc_CO2_gas_flux)
        return s

    def c_output(self):
        s = abi.BgcOutput()
        s.BGC_tendencies = abi.raw_dptr(self.BGC_tendencies.data_ptr())
        s.PH_PREV_3D = abi.raw_dptr(self.PH_PREV_3D.data_ptr())
        s.PH_PREV_ALT_CO2_3D = abi.raw_dptr(self.PH_PREV_ALT_CO2_3D.data_ptr())
        return s

    def c_diag(self, enabled=True):
        s = abi.BgcDiagnostics()
        if enabled:
            for n, t in self.diag.items():
                setattr(s, n, abi.raw_dptr(t.data_ptr()))
        return s

    def c_flux_diag(self):
        s = abi.BgcFluxDiagnostics()
        for n, t in self.flux_diag.items():
            setattr(s, n, abi.raw_dptr(t.data_ptr()))
        return s


class DeviceDmsColumns(_DeviceMixin):
    def __init__(self, nLevelsMax, nColumnsMax, nColumns=None, device="cuda:0", diagnostics=True):
        torch = _torch()
        self.device = torch.device(device)
        nL, nC = int(nLevelsMax), int(nColumnsMax)
        self.nLevelsMax, self.nColumnsMax = nL, nC
        self.nColumns = nC if nColumns is None else int(nColumns)
        A = self._alloc
        self.DMS_tracers = A((abi.DMS_TRACER_CNT, nL, nC))
        self.cell_thickness = A((nL, nC))
        self.number_of_active_levels = A((nC,), torch.int32)
        self.forcing = {n: A((nC,)) for n in abi.DMS_FORCING_C1}
        self.forcing["netFlux"] = A((abi.DMS_TRACER_CNT, nC))
        self.lcalc_DMS_gas_flux = 1
        self.DMS_tendencies = A((abi.DMS_TRACER_CNT, nL, nC))
        self.diag = {n: A((nL, nC)) for n in abi.DMS_DIAG} if diagnostics else {}
        self.flux_diag = {n: A((nC,)) for n in abi.DMS_FLUX_DIAG}

    def nbytes(self):
        t = [self.DMS_tracers, self.cell_thickness, self.number_of_active_levels, self.DMS_tendencies]
        t += list(self.forcing.values()) + list(self.diag.values()) + list(self.flux_diag.values())
        return sum(x.numel() * x.element_size() for x in t)

    def load(self, host):
        torch = _torch()

        def put(dst, arr):
            dst.copy_(torch.from_numpy(arr))
        put(self.DMS_tracers, self._soa(host.DMS_tracers))
        put(self.cell_thickness, self._soa(host.cell_thickness))
        put(self.number_of_active_levels, host.number_of_active_levels)
        for n, t in self.forcing.items():
            put(t, self._flux(host.forcing[n]))
        put(self.DMS_tendencies, self._soa(host.DMS_tendencies))
        for n, t in self.diag.items():
            put(t, self._soa(host.diag[n]))
        for n, t in self.flux_diag.items():
            put(t, host.flux_diag[n])
        self.nColumns = host.nColumns
        self.lcalc_DMS_gas_flux = host.lcalc_DMS_gas_flux
        _sync(self.device)   # torch copies run on torch's stream, library calls on the ctx stream
        return self

    def store(self, host):
        host.DMS_tendencies[...] = np.transpose(self.DMS_tendencies.cpu().numpy(), (1, 2, 0))
        for n, t in self.diag.items():
            host.diag[n][...] = t.cpu().numpy()
        for n, t in self.flux_diag.items():
            host.flux_diag[n][...] = t.cpu().numpy()
        for n, t in self.forcing.items():
            a = t.cpu().numpy()
            host.forcing[n][...] = a.T if a.ndim == 2 else a
        return host

    def c_input(self):
        s = abi.DmsInput()
        s.DMS_tracers = abi.raw_dptr(self.DMS_tracers.data_ptr())
        s.cell_thickness = abi.raw_dptr(self.cell_thickness.data_ptr())
        s.number_of_active_levels = abi.raw_iptr(self.number_of_active_levels.data_ptr())
        return s

    def c_forcing(self):
        s = abi.DmsForcing()
        for n, t in self.forcing.items():
            setattr(s, n, abi.raw_dptr(t.data_ptr()))
        s.lcalc_DMS_gas_flux = int(self.lcalc_DMS_gas_flux)
        return s

    def c_output(self):
        s = abi.DmsOutput()
        s.DMS_tendencies = abi.raw_dptr(self.DMS_tendencies.data_ptr())
        return s

    def c_diag(self, enabled=True):
        s = abi.DmsDiagnostics()
        if enabled:
            for n, t in self.diag.items():
                setattr(s, n, abi.raw_dptr(t.data_ptr()))
        return s

    def c_flux_diag(self):
        s = abi.DmsFluxDiagnostics()
        for n, t in self.flux_diag.items():
            setattr(s, n, abi.raw_dptr(t.data_ptr()))
        return s


class DeviceMacrosColumns(_DeviceMixin):
    def __init__(self, nLevelsMax, nColumnsMax, nColumns=None, device="cuda:0", diagnostics=True):
        torch = _torch()
        self.device = torch.device(device)
        nL, nC = int(nLevelsMax), int(nColumnsMax)
        self.nLevelsMax, self.nColumnsMax = nL, nC
        self.nColumns = nC if nColumns is None else int(nColumns)
        A = self._alloc
        self.MACROS_tracers = A((abi.MACROS_TRACER_CNT, nL, nC))
        self.cell_thickness = A((nL, nC))
        self.number_of_active_levels = A((nC,), torch.int32)
        self.MACROS_tendencies = A((abi.MACROS_TRACER_CNT, nL, nC))
        self.diag = {n: A((nL, nC)) for n in abi.MACROS_DIAG} if diagnostics else {}

    def nbytes(self):
        t = [self.MACROS_tracers, self.cell_thickness, self.number_of_active_levels, self.MACROS_tendencies]
        t += list(self.diag.values())
        return sum(x.numel() * x.element_size() for x in t)

    def load(self, host):
        torch = _torch()

        def put(dst, arr):
            dst.copy_(torch.from_numpy(arr))
        put(self.MACROS_tracers, self._soa(host.MACROS_tracers))
        put(self.cell_thickness, self._soa(host.cell_thickness))
        put(self.number_of_active_levels, host.number_of_active_levels)
        put(self.MACROS_tendencies, self._soa(host.MACROS_tendencies))
        for n, t in self.diag.items():
            put(t, self._soa(host.diag[n]))
        self.nColumns = host.nColumns
        _sync(self.device)   # torch copies run on torch's stream, library calls on the ctx stream
        return self

    def store(self, host):
        host.MACROS_tendencies[...] = np.transpose(self.MACROS_tendencies.cpu().numpy(), (1, 2, 0))
        for n, t in self.diag.items():
            host.diag[n][...] = t.cpu().numpy()
        return host

    def c_input(self):
        s = abi.MacrosInput()
        s.MACROS_tracers = abi.raw_dptr(self.MACROS_tracers.data_ptr())
        s.cell_thickness = abi.raw_dptr(self.cell_thickness.data_ptr())
        s.number_of_active_levels = abi.raw_iptr(self.number_of_active_levels.data_ptr())
        return s

    def c_output(self):
        s = abi.MacrosOutput()
        s.MACROS_tendencies = abi.raw_dptr(self.MACROS_tendencies.data_ptr())
        return s

    def c_diag(self, enabled=True):
        s = abi.MacrosDiagnostics()
        if enabled:
            for n, t in self.diag.items():
                setattr(s, n, abi.raw_dptr(t.data_ptr()))
        return s


__all__ = ["BgcError", "Parms", "BGC_parms_init", "Context", "BGC_SourceSink", "BGC_SurfaceFluxes",
           "DMS_SourceSink", "DMS_SurfaceFluxes", "MACROS_SourceSink", "co2calc_points",
           "comp_CO3terms_points", "comp_co3_sat_vals_points",
           "co2calc_points_device", "DeviceBgcColumns", "DeviceDmsColumns", "DeviceMacrosColumns",
           "BgcColumns", "DmsColumns", "MacrosColumns", "lib", "ABI_SYMBOLS"]
