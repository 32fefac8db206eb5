"""ctypes mirror of include/bgc_b200.h.

The header is the single source of truth: the struct layouts and the X-macro
field lists are parsed from it at import time, so the Python host layer, the
tests and the Fortran shim documentation can never drift from the C ABI.
(Reference types being mirrored: BGC_parms.F90:51-321, DMS_parms.F90:62-154,
MACROS_parms.F90:62-113.)
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
HEADER = os.path.join(REPO, "include", "bgc_b200.h")


def _strip_comments(src):
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)


def _parse_header(path):
    raw = open(path).read()
    src = _strip_comments(raw).replace("\\\n", " ")
    defines, lists = {}, {}
    for m in re.finditer(r"#define\s+(\w+)\(X\)\s+(.*)", src):
        lists[m.group(1)] = re.findall(r"X\((\w+)\)", m.group(2))
    for m in re.finditer(r"#define\s+(\w+)\s+(-?\d+)\s*$", src, flags=re.M):
        defines[m.group(1)] = int(m.group(2))
    for m in re.finditer(r"enum\s*\{(.*?)\}", src, flags=re.S):
        nxt = 0
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = item.split("=")
                nxt = int(v.strip())
                defines[k.strip()] = nxt
            else:
                defines[item] = nxt
            nxt += 1
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        name, body = m.group(3), m.group(2)
        # expand X-macro lists:  LIST(BGC_DECL_PTR) -> double *a; double *b; ...
        def expand(mm):
            return " ".join("double *%s;" % f for f in lists[mm.group(1)])
        body = re.sub(r"(\w+_LIST)\(BGC_DECL_PTR\)", expand, body)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            mm = re.match(r"(const\s+)?(unsigned\s+long\s+long|double|int)\s+(.*)", decl, flags=re.S)
            if not mm:
                raise ValueError("cannot parse declaration %r in struct %s" % (decl, name))
            base = {"double": C.c_double, "int": C.c_int,
                    "unsigned long long": C.c_ulonglong}[re.sub(r"\s+", " ", mm.group(2))]
            for item in mm.group(3).split(","):
                item = item.strip()
                ptr = item.startswith("*")
                item = item.lstrip("* ").strip()
                arr = re.match(r"(\w+)\[(\w+)\]", item)
                if arr:
                    n = arr.group(2)
                    n = int(n) if n.isdigit() else defines[n]
                    fields.append((arr.group(1), base * n))
                elif ptr:
                    fields.append((item, C.POINTER(base)))
                else:
                    fields.append((item, base))
        structs[name] = fields
    return defines, lists, structs


DEFINES, LISTS, _STRUCT_FIELDS = _parse_header(HEADER)

BGC_TRACER_CNT = DEFINES["BGC_TRACER_CNT"]
DMS_TRACER_CNT = DEFINES["DMS_TRACER_CNT"]
MACROS_TRACER_CNT = DEFINES["MACROS_TRACER_CNT"]
BGC_AUTOTROPH_CNT = DEFINES["BGC_AUTOTROPH_CNT"]
BGC_INVENTORY_LEN = DEFINES["BGC_INVENTORY_LEN"]
BGC_MEM_HOST_FORTRAN = DEFINES["BGC_MEM_HOST_FORTRAN"]
BGC_MEM_DEVICE_SOA = DEFINES["BGC_MEM_DEVICE_SOA"]
BGC_OK = DEFINES["BGC_OK"]


def _mk(name):
    return type(name, (C.Structure,), {"_fields_": _STRUCT_FIELDS[name]})


BgcParams = _mk("BgcParams")
BgcAutotroph = _mk("BgcAutotroph")
BgcIndices = _mk("BgcIndices")
DmsParams = _mk("DmsParams")
DmsIndices = _mk("DmsIndices")
MacrosParams = _mk("MacrosParams")
MacrosIndices = _mk("MacrosIndices")
BgcInput = _mk("BgcInput")
BgcForcing = _mk("BgcForcing")
BgcOutput = _mk("BgcOutput")
BgcFluxDiagnostics = _mk("BgcFluxDiagnostics")
BgcDiagnostics = _mk("BgcDiagnostics")
DmsInput = _mk("DmsInput")
DmsForcing = _mk("DmsForcing")
DmsOutput = _mk("DmsOutput")
DmsFluxDiagnostics = _mk("DmsFluxDiagnostics")
DmsDiagnostics = _mk("DmsDiagnostics")
MacrosInput = _mk("MacrosInput")
MacrosOutput = _mk("MacrosOutput")
MacrosDiagnostics = _mk("MacrosDiagnostics")
BgcStatus = _mk("BgcStatus")
BgcAutotroph4 = BgcAutotroph * BGC_AUTOTROPH_CNT

# diagnostic field lists by shape class (see header)
BGC_DIAG_K2 = LISTS["BGC_DIAG_K2_LIST"]
BGC_DIAG_KA = LISTS["BGC_DIAG_KA_LIST"]
BGC_DIAG_CA = LISTS["BGC_DIAG_CA_LIST"]
BGC_DIAG_C1 = LISTS["BGC_DIAG_C1_LIST"]
BGC_FLUX_DIAG = LISTS["BGC_FLUX_DIAG_LIST"]
DMS_DIAG = LISTS["DMS_DIAG_LIST"]
DMS_FLUX_DIAG = LISTS["DMS_FLUX_DIAG_LIST"]
MACROS_DIAG = LISTS["MACROS_DIAG_LIST"]
# declared in BGC_diagnostics_type but never zeroed nor written by the reference
BGC_DIAG_UNTOUCHED = ("diag_POC_ACCUM", "diag_DONr_remin", "diag_DOPr_remin")

BGC_TRACER_NAMES = [f[0][:-4] for f in _STRUCT_FIELDS["BgcIndices"][:BGC_TRACER_CNT]]
DMS_TRACER_NAMES = [f[0][:-4] for f in _STRUCT_FIELDS["DmsIndices"]]
MACROS_TRACER_NAMES = [f[0][:-4] for f in _STRUCT_FIELDS["MacrosIndices"]]

BGC_FORCING_K2 = ["FESEDFLUX", "NUTR_RESTORE_RTAU", "NO3_CLIM", "PO4_CLIM", "SiO3_CLIM"]
BGC_FORCING_C1 = ["dust_FLUX_IN", "ShortWaveFlux_surface", "surfacePressure", "iceFraction",
                  "windSpeedSquared10m", "atmCO2", "atmCO2_ALT_CO2", "surface_pH",
                  "surface_pH_alt_co2", "surfaceDepth", "SST", "SSS"]
BGC_FORCING_FLUX = ["depositionFlux", "riverFlux", "gasFlux", "seaIceFlux", "netFlux"]
DMS_FORCING_C1 = ["ShortWaveFlux_surface", "surfacePressure", "iceFraction",
                  "windSpeedSquared10m", "SST", "SSS"]


def struct_to_dict(s):
    out = {}
    for name, typ in s._fields_:
        v = getattr(s, name)
        out[name] = list(v) if hasattr(v, "__len__") else v
    return out


def dptr(a):
    """double* view of a numpy float64 array (or None -> NULL)."""
    if a is None:
        return C.POINTER(C.c_double)()
    assert a.dtype.name == "float64" and (a.flags["F_CONTIGUOUS"] or a.flags["C_CONTIGUOUS"])
    return a.ctypes.data_as(C.POINTER(C.c_double))


def fptr(a):
    """double* of an argument of the Fortran-layout API: arrays of rank > 1 must be column-major
    (a C-ordered (k,col) array would silently be read transposed)."""
    if a is not None and a.ndim > 1:
        assert a.flags["F_CONTIGUOUS"], "Fortran-layout (column-major) array required"
    return dptr(a)


def iptr(a):
    assert a.dtype.name == "int32"
    return a.ctypes.data_as(C.POINTER(C.c_int))


def raw_dptr(addr):
    """double* from a raw (device) address."""
    return C.cast(C.c_void_p(int(addr)), C.POINTER(C.c_double))


def raw_iptr(addr):
    return C.cast(C.c_void_p(int(addr)), C.POINTER(C.c_int))
