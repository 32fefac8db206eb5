"""ctypes binding of oracle/_ref/libbgc_ref.so — the UNMODIFIED reference Fortran
(/root/reference/*.F90) machine-translated to C by oracle/f90c.py and compiled by gcc.

TEST INFRASTRUCTURE ONLY.  This is what pins the hand-written oracle: the same inputs go
through the translated reference and through oracle/libbgc_oracle.so and the results are
compared bit for bit (tests/test_reference_translated.py); tests/golden/*.npz are generated
from the translated reference (tests/golden/make_golden.py).  The library and the generated C
live in oracle/_ref/ (git-ignored: nothing derived from the reference sources is committed);
`build()` regenerates them wherever /root/reference exists, and the prebuilt files travel to
the GPU box with the snapshot.

The reference keeps state in module variables; the translation makes them thread-local, so a
thread that wants to compute must run `RefParms()` (the *_parms_init / *_init calls) itself.
"""
import ctypes as C
import json
import os
import subprocess
import sys
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")
LIB_PATH = os.environ.get("BGC_REF_LIB", os.path.join(REFDIR, "libbgc_ref.so"))   # override: coverage / profile builds
META_PATH = os.path.join(REFDIR, "meta.json")
REFERENCE_SRC = os.environ.get("BGC_REFERENCE_SRC", "/root/reference")

class FA(C.Structure):
    """array descriptor of the translation: base pointer + up to three extents"""
    _fields_ = [("p", C.c_void_p), ("n1", C.c_int), ("n2", C.c_int), ("n3", C.c_int)]


_SCALAR = {"i1": C.c_byte, "i2": C.c_short, "i4": C.c_int, "i8": C.c_longlong, "r4": C.c_float,
           "r8": C.c_double, "log": C.c_int, "cptr": C.c_void_p}


class TLib:
    """One translated library (shared object + the translator's metadata)."""

    def __init__(self, lib_path, meta_path, preload=()):
        self.lib_path, self.meta_path, self.preload = lib_path, meta_path, list(preload)
        self._lib = self._meta = None
        self._structs = {}
        self._pristine = {}

    def available(self):
        return os.path.exists(self.lib_path) and os.path.exists(self.meta_path)

    def ctype(self, code):
        if code in _SCALAR:
            return _SCALAR[code]
        if code.startswith("char"):
            return C.c_char * int(code[4:])
        if code.startswith("type:"):
            return self.struct(code[5:])
        raise KeyError(code)

    def field_ctype(self, f):
        if f["alloc"]:
            return FA
        t = self.ctype(f["type"])
        if f["rank"] > 0:
            t = t * int(np.prod(f["dims"]))
        return t

    def meta(self):
        if self._meta is None:
            self._meta = json.load(open(self.meta_path))
        return self._meta

    def struct(self, cname):
        """ctypes mirror of a translated derived type (fields in declaration order)."""
        if cname not in self._structs:
            fields = [(f["cname"], self.field_ctype(f)) for f in self.meta()["types"][cname]]
            self._structs[cname] = type(cname, (C.Structure,), {"_fields_": fields})
        return self._structs[cname]

    def lib(self):
        if self._lib is None:
            if not self.available():
                if not can_build():
                    raise RuntimeError(self.lib_path + " is missing and the reference sources are not "
                                       "available to rebuild it")
                build()
            self._keep = [C.CDLL(p, mode=C.RTLD_GLOBAL) for p in self.preload]
            self._lib = C.CDLL(self.lib_path)
        return self._lib

    def const(self, cname):
        """value of a module-level named constant (e.g. 'bgc_parms__epsc')"""
        f = getattr(self.lib(), "ref_const__" + cname)
        f.restype = self.ctype(self.meta()["consts"][cname])
        return f()

    def var(self, cname):
        """ctypes object aliasing a module variable OF THE CALLING THREAD"""
        f = getattr(self.lib(), "ref_addr__" + cname)
        f.restype = C.c_void_p
        return self.field_ctype(self.meta()["vars"][cname]).from_address(f())

    # module variables a host sets like a namelist: the *_parms tunables and BGC_mod's restoring switches
    _TUNABLE_PREFIXES = ("bgc_parms__", "dms_parms__", "macros_parms__", "bgc_mod__lrest_")

    def restore_tunables(self):
        """Put the tunable module variables of the calling thread back to what they were when this
        thread first asked.  `*_parms_init` does not assign all of them: `f_qsw_par_DMS` gets its value
        in its declaration (DMS_parms.F90:191-192) and the `lrest_*` switches are plain module variables
        of BGC_mod (BGC_mod.F90:131-134), so what an earlier caller stored there (RefParms.sync_from with
        perturbed tables) would otherwise reach every later caller of the process - as it would in the
        Fortran; a test that asks for the defaults means the defaults."""
        key = threading.get_ident()
        names = [cn for cn, info in self.meta()["vars"].items()
                 if cn.startswith(self._TUNABLE_PREFIXES) and not info["alloc"]]
        snap = self._pristine.get(key)
        if snap is None:
            snap = {}
            for cn in names:
                v = self.var(cn)
                snap[cn] = C.string_at(C.addressof(v), C.sizeof(v))
            self._pristine[key] = snap
            return
        for cn, raw in snap.items():
            C.memmove(C.addressof(self.var(cn)), raw, len(raw))

    def call(self, cname, *args):
        """Call a translated procedure.  Scalars may be given as Python numbers (wrapped, passed
        by reference; the ctypes objects are returned so that intent(out) values can be read),
        derived types as ctypes Structures / arrays of them."""
        pm = self.meta()["procs"][cname]
        fn = getattr(self.lib(), cname)
        fn.restype = self.ctype(pm["result"]) if pm["result"] else None
        if len(args) != len(pm["args"]):
            raise TypeError(f"{cname}: {len(args)} arguments for {len(pm['args'])} dummies")
        boxed = []
        for a, am in zip(args, pm["args"]):
            if isinstance(a, (C.Structure, C.Array, C._SimpleCData)):
                boxed.append(a)
            else:
                boxed.append(self.ctype(am["type"])(a))
        r = fn(*[C.byref(b) for b in boxed])
        return r, boxed

    def fill(self, cname, arrays, scalars=None):
        """Build a translated derived type whose allocatable components alias the numpy arrays
        in `arrays` (component names compared case-insensitively) and whose scalar components
        come from `scalars`.  Returns (struct, keepalive)."""
        s = self.struct(cname)()
        low = {k.lower(): v for k, v in arrays.items()}
        sc = {k.lower(): v for k, v in (scalars or {}).items()}
        keep = []
        for f in self.meta()["types"][cname]:
            n = f["name"]
            if f["alloc"]:
                if n in low and low[n] is not None:
                    a = low[n]
                    if f["type"].startswith("r8"):
                        assert a.dtype == np.float64, n
                    elif f["type"] == "i4":
                        assert a.dtype == np.int32, n
                    assert a.ndim == f["rank"], (n, a.shape, f["rank"])
                    setattr(s, f["cname"], describe(a))
                    keep.append(a)
            elif n in sc:
                setattr(s, f["cname"], sc[n])
        return s, keep


REF = TLib(LIB_PATH, META_PATH)


def available():
    return REF.available()


def can_build():
    return os.path.exists(os.path.join(REFERENCE_SRC, "BGC_mod.F90"))


def build():
    """Translate + compile (needs the reference sources; a no-op when up to date)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "ref", f"REFERENCE_SRC={REFERENCE_SRC}"])


def meta():
    return REF.meta()


def struct(cname):
    return REF.struct(cname)


def lib():
    return REF.lib()


def const(cname):
    return REF.const(cname)


def var(cname):
    return REF.var(cname)


def call(cname, *args):
    return REF.call(cname, *args)


def fill(cname, arrays, scalars=None):
    return REF.fill(cname, arrays, scalars)


def describe(arr):
    """descriptor of a Fortran-ordered numpy array (no copy: the routine works in place)"""
    assert arr.flags["F_CONTIGUOUS"] or arr.ndim == 1
    d = FA()
    d.p = arr.ctypes.data
    sh = list(arr.shape) + [1, 1, 1]
    d.n1, d.n2, d.n3 = sh[0], sh[1], sh[2]
    return d


def _arrays_of(cols):
    out = {}
    for k, v in vars(cols).items():
        if isinstance(v, np.ndarray):
            out[k] = v
        elif isinstance(v, dict):
            for kk, vv in v.items():
                if isinstance(vv, np.ndarray):
                    out[kk] = vv
    return out


def _scalars_of(cols):
    return {k: int(v) for k, v in vars(cols).items() if k.startswith("lcalc_")}


class RefParms:
    """BGC_parms_init + BGC_init + DMS_parms_init + DMS_init + MACROS_parms_init + MACROS_init
    of the translated reference, with the host-chosen tracer slots of `parms` (an
    oracle.Parms) and the host-set T0_Kelvin_BGC (quirk Q7)."""

    def __init__(self, parms=None, t0_kelvin=273.15, L=None):
        self.L = L = L or REF
        m, struct, var, call = L.meta(), L.struct, L.var, L.call
        L.restore_tunables()      # defaults mean defaults, whatever an earlier sync_from left behind
        self.ind = struct("bgc_parms__bgc_indices_type")()
        self.autotrophs = (struct("bgc_parms__autotroph_type") * 4)()
        self._names = {}
        for tname, obj_name, n in (("bgc_parms__bgc_indices_type", "ind", 30),
                                   ("dms_parms__dms_indices_type", "dms_ind", 14),
                                   ("macros_parms__macros_indices_type", "macros_ind", 8)):
            obj = struct(tname)() if obj_name != "ind" else self.ind
            setattr(self, obj_name, obj)
            for f in m["types"][tname]:
                if f["alloc"] and f["type"].startswith("char"):
                    buf = np.full((n, int(f["type"][4:])), ord(" "), dtype=np.uint8)
                    d = FA()
                    d.p, d.n1, d.n2, d.n3 = buf.ctypes.data, n, 1, 1
                    setattr(obj, f["cname"], d)
                    self._names[(obj_name, f["name"])] = buf
        # the host (MPAS) assigns the tracer slots before calling *_init
        if parms is not None:
            self._copy_slots(parms.ind, self.ind, 30)
            self._copy_slots(parms.dms_ind, self.dms_ind, None)
            self._copy_slots(parms.macros_ind, self.macros_ind, None)
        else:      # declaration order 1..N (the autotroph indices sp_ind.. are set by BGC_parms_init)
            for obj, count in ((self.ind, 30), (self.dms_ind, 14), (self.macros_ind, 8)):
                slots = [n for n, t in obj._fields_ if t is C.c_int and n.endswith("_ind")][:count]
                for i, fname in enumerate(slots):
                    setattr(obj, fname, i + 1)
        var("bgc_parms__t0_kelvin_bgc").value = t0_kelvin
        call("bgc_parms__bgc_parms_init", self.ind, self.autotrophs)
        call("bgc_mod__bgc_init", self.ind, self.autotrophs)
        call("dms_parms__dms_parms_init")
        call("dms_mod__dms_init", self.dms_ind)
        call("macros_parms__macros_parms_init")
        call("macros_mod__macros_init", self.macros_ind)

    @staticmethod
    def _copy_slots(src, dst, limit):
        names = [n for n, _ in src._fields_]
        if limit is not None:
            names = names[:limit]
        for n in names:
            setattr(dst, n.lower(), getattr(src, n))

    def sync_from(self, parms):
        """Copy the run-time tunables and the functional-group table of an oracle.Parms /
        host.Parms into the module variables of the translated reference (what a namelist read
        does upstream).  Compile-time constants (epsC ...) are not settable and are skipped."""
        m, var = self.L.meta(), self.L.var
        for src, mod in ((parms.bgc, "bgc_parms"), (parms.dms, "dms_parms"), (parms.macros, "macros_parms")):
            for n, _ in src._fields_:
                cn = f"{mod}__{n.lower()}"
                if n.startswith("lrest_"):
                    cn = f"bgc_mod__{n.lower()}"
                if cn not in m["vars"]:
                    continue
                v, dst = getattr(src, n), var(cn)
                if hasattr(v, "__len__"):
                    for i, x in enumerate(v):
                        dst[i] = x
                else:
                    dst.value = v
        for i in range(4):
            for n, _ in parms.autotrophs[i]._fields_:
                setattr(self.autotrophs[i], n.lower(), getattr(parms.autotrophs[i], n))
        return self

    def name(self, which, field, i):
        return bytes(self._names[(which, field)][i]).decode().rstrip()


def BGC_SourceSink(rp, cols, alt_co2_use_eco=True):
    arrs = _arrays_of(cols)
    fill, call = rp.L.fill, rp.L.call
    cin, k1 = fill("bgc_parms__bgc_input_type", arrs)
    cfo, k2 = fill("bgc_parms__bgc_forcing_type", arrs, _scalars_of(cols))
    cout, k3 = fill("bgc_parms__bgc_output_type", arrs)
    cdg, k4 = fill("bgc_parms__bgc_diagnostics_type", arrs)
    call("bgc_mod__bgc_sourcesink", rp.autotrophs, rp.ind, cin, cfo, cout, cdg,
         cols.nLevelsMax, cols.nColumnsMax, cols.nColumns, int(alt_co2_use_eco))


def BGC_SurfaceFluxes(rp, cols):
    arrs = _arrays_of(cols)
    fill, call = rp.L.fill, rp.L.call
    cin, k1 = fill("bgc_parms__bgc_input_type", arrs)
    cfo, k2 = fill("bgc_parms__bgc_forcing_type", arrs, _scalars_of(cols))
    cfd, k3 = fill("bgc_parms__bgc_flux_diagnostics_type", arrs)
    call("bgc_mod__bgc_surfacefluxes", rp.ind, cin, cfo, cfd, cols.nColumnsMax, cols.nColumns)


def DMS_SourceSink(rp, cols):
    arrs = _arrays_of(cols)
    fill, call = rp.L.fill, rp.L.call
    cin, k1 = fill("dms_parms__dms_input_type", arrs)
    cfo, k2 = fill("dms_parms__dms_forcing_type", arrs, _scalars_of(cols))
    cout, k3 = fill("dms_parms__dms_output_type", arrs)
    cdg, k4 = fill("dms_parms__dms_diagnostics_type", arrs)
    call("dms_mod__dms_sourcesink", rp.dms_ind, cin, cfo, cout, cdg,
         cols.nLevelsMax, cols.nColumnsMax, cols.nColumns)


def DMS_SurfaceFluxes(rp, cols):
    arrs = _arrays_of(cols)
    fill, call = rp.L.fill, rp.L.call
    cin, k1 = fill("dms_parms__dms_input_type", arrs)
    cfo, k2 = fill("dms_parms__dms_forcing_type", arrs, _scalars_of(cols))
    cfd, k3 = fill("dms_parms__dms_flux_diagnostics_type", arrs)
    call("dms_mod__dms_surfacefluxes", rp.dms_ind, cin, cfo, cfd, cols.nColumnsMax, cols.nColumns)


def MACROS_SourceSink(rp, cols):
    arrs = _arrays_of(cols)
    fill, call = rp.L.fill, rp.L.call
    cin, k1 = fill("macros_parms__macros_input_type", arrs)
    cout, k2 = fill("macros_parms__macros_output_type", arrs)
    cdg, k3 = fill("macros_parms__macros_diagnostics_type", arrs)
    call("macros_mod__macros_sourcesink", rp.macros_ind, cin, cout, cdg,
         cols.nLevelsMax, cols.nColumnsMax, cols.nColumns)


def co2calc_1point(depth, temp, salt, dic, ta, pt, sit, phlo, phhi, xco2, atmpres,
                   locmip_k1_k2_bug_fix=True, lcomp_co3_coeffs=True):
    _, b = call("co2calc__co2calc_1point", depth, int(locmip_k1_k2_bug_fix), int(lcomp_co3_coeffs),
                temp, salt, dic, ta, pt, sit, phlo, phhi, 0.0, xco2, atmpres, 0.0, 0.0, 0.0, 0.0)
    return dict(phlo=b[9].value, phhi=b[10].value, ph=b[11].value, co2star=b[14].value,
                dco2star=b[15].value, pco2surf=b[16].value, dpco2=b[17].value)


def comp_CO3terms(k, depth, temp, salt, dic, ta, pt, sit, phlo, phhi, lcomp_co3_coeffs=True):
    _, b = call("co2calc__comp_co3terms", int(k), depth, int(lcomp_co3_coeffs), temp, salt, dic, ta,
                pt, sit, phlo, phhi, 0.0, 0.0, 0.0, 0.0)
    return dict(phlo=b[9].value, phhi=b[10].value, pH=b[11].value, H2CO3=b[12].value,
                HCO3=b[13].value, CO3=b[14].value)


def comp_co3_sat_vals(k, depth, temp, salt):
    _, b = call("co2calc__comp_co3_sat_vals", int(k), depth, temp, salt, 0.0, 0.0)
    return b[4].value, b[5].value


# ------------------------------------------------------------------ threaded driver (bench.py)

class SlabRunner:
    """Runs the translated reference on column slabs from a pool of threads.  The reference is
    serial and not re-entrant (module SAVE state); the translation makes that state thread-local,
    so every worker thread initialises its own copy (RefParms) once and then works through slabs.
    ctypes releases the GIL for the duration of each call."""

    def __init__(self, parms, nthreads):
        import threading
        from concurrent.futures import ThreadPoolExecutor
        self.parms, self.nthreads = parms, int(nthreads)
        self._tls = threading.local()
        self._pool = ThreadPoolExecutor(max_workers=self.nthreads)

    def _rp(self):
        if not hasattr(self._tls, "rp"):
            self._tls.rp = RefParms(self.parms)
        return self._tls.rp

    def _one(self, slab):
        rp = self._rp()
        bgc, dms, mac = slab
        BGC_SourceSink(rp, bgc, True)
        BGC_SurfaceFluxes(rp, bgc)
        if dms is not None:
            DMS_SourceSink(rp, dms)
            DMS_SurfaceFluxes(rp, dms)
        if mac is not None:
            MACROS_SourceSink(rp, mac)

    def step(self, slabs):
        """one BGC + DMS + MACROS step over every (bgc, dms, macros) slab"""
        list(self._pool.map(self._one, slabs))

    def close(self):
        self._pool.shutdown()


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    print("available:", available(), LIB_PATH)
