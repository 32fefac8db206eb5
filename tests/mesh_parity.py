"""Cell-for-cell comparison of a DEVICE-RESIDENT mesh (one C-ABI call per procedure over the
whole mesh) with the translated reference (oracle/_ref/libbgc_ref.so), at mesh sizes where a
host copy of every array would not fit comfortably: the mesh is cut into units of a few
hundred columns, the reference runs unit by unit on a pool of host threads, each unit's GPU
results are fetched from the device containers and compared, and the per-array statistics
(max |gpu - ref| and max |ref|) are folded over the units.  The verdict is therefore the
same whole-array criterion as tests/parity.py (max|gpu-ref| / max|ref| per array), at the
same tolerances.  Test infrastructure only.
"""
import os
import sys
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

import parity

pkg = parity.pkg
abi = pkg.abi
sys.path.insert(0, os.path.join(parity.REPO, "oracle"))
import ref_translated as rt   # noqa: E402  (test infrastructure only)


class FieldStats:
    """per field: max |got - ref|, max |ref| (or an explicit scale), non-finite mismatches"""

    def __init__(self):
        self.d, self.m, self.bad = {}, {}, {}
        self._lock = threading.Lock()

    def add(self, name, got, ref, scale=None):
        got = np.asarray(got, dtype=np.float64)
        ref = np.asarray(ref, dtype=np.float64)
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        fin = np.isfinite(ref)
        bad = int((~np.isfinite(got) & fin).sum())
        d = float(np.max(np.abs(np.where(fin, got - ref, 0.0)))) if got.size else 0.0
        if not np.isfinite(d):
            d, bad = 0.0, bad + 1
        m = float(np.max(np.abs(np.where(fin, ref, 0.0)))) if ref.size else 0.0
        if scale is not None:
            m = max(m, float(scale))
        with self._lock:
            self.d[name] = max(self.d.get(name, 0.0), d)
            self.m[name] = max(self.m.get(name, 0.0), m)
            self.bad[name] = self.bad.get(name, 0) + bad

    def errors(self):
        out = {}
        for k in self.d:
            if self.bad[k]:
                out[k] = float("inf")
            elif self.m[k] == 0.0:
                out[k] = 0.0 if self.d[k] == 0.0 else float("inf")
            else:
                out[k] = self.d[k] / self.m[k]
        return out

    def check(self, what):
        errs = self.errors()
        fails = []
        for k, e in errs.items():
            base = k.split(":", 1)[-1]
            solver = (base in parity.SOLVER_DIAGS or base.startswith("PH_PREV") or base in parity.SOLVER_FLUX
                      or base in ("gasFlux", "netFlux", "surface_pH", "surface_pH_alt_co2"))
            lim = parity.TOL_SOLVER if solver else parity.TOL_TEND
            if not (e <= lim):
                fails.append("%s: %.3e > %.1e" % (k, e, lim))
        if fails:
            raise AssertionError("%s: parity failed for %d of %d arrays:\n  %s"
                                 % (what, len(fails), len(errs), "\n  ".join(fails[:40])))
        return errs


# ---------------------------------------------------------------- device slab -> host container
def _get(t, like, sl):
    """device tensor ([n,] [k,] col) restricted to columns `sl` -> host array `like` ((k, col, n) / (k, col) / (col))"""
    a = t[..., sl].cpu().numpy()
    like[...] = np.transpose(a, (1, 2, 0)) if like.ndim == 3 else a


def fetch_bgc(dev, host_cols, sl):
    _get(dev.BGC_tendencies, host_cols.BGC_tendencies, sl)
    _get(dev.PH_PREV_3D, host_cols.PH_PREV_3D, sl)
    _get(dev.PH_PREV_ALT_CO2_3D, host_cols.PH_PREV_ALT_CO2_3D, sl)
    for n, t in dev.diag.items():
        if n in abi.BGC_DIAG_CA:
            host_cols.diag[n][...] = t[:, sl].cpu().numpy().T
        else:
            _get(t, host_cols.diag[n], sl)
    for n, t in dev.flux_diag.items():
        host_cols.flux_diag[n][...] = t[sl].cpu().numpy()
    for n, t in dev.forcing.items():
        if n in abi.BGC_FORCING_FLUX:
            host_cols.forcing[n][...] = t[:, sl].cpu().numpy().T
        else:
            _get(t, host_cols.forcing[n], sl)


def fetch_dms(dev, host_cols, sl):
    host_cols.DMS_tendencies[...] = np.transpose(dev.DMS_tendencies[..., sl].cpu().numpy(), (1, 2, 0))
    for n, t in dev.diag.items():
        host_cols.diag[n][...] = t[:, sl].cpu().numpy()
    for n, t in dev.flux_diag.items():
        host_cols.flux_diag[n][...] = t[sl].cpu().numpy()
    for n, t in dev.forcing.items():
        a = t[..., sl].cpu().numpy()
        host_cols.forcing[n][...] = a.T if a.ndim == 2 else a


def fetch_macros(dev, host_cols, sl):
    host_cols.MACROS_tendencies[...] = np.transpose(dev.MACROS_tendencies[..., sl].cpu().numpy(), (1, 2, 0))
    for n, t in dev.diag.items():
        host_cols.diag[n][...] = t[:, sl].cpu().numpy()


# ---------------------------------------------------------------- accumulate the comparisons
def add_bgc(st, tag, ref, got, surface):
    for n in range(abi.BGC_TRACER_CNT):
        st.add("%s:tend[%d]" % (tag, n + 1), got.BGC_tendencies[:, :, n], ref.BGC_tendencies[:, :, n])
    st.add(tag + ":PH_PREV_3D", got.PH_PREV_3D, ref.PH_PREV_3D)
    st.add(tag + ":PH_PREV_ALT_CO2_3D", got.PH_PREV_ALT_CO2_3D, ref.PH_PREV_ALT_CO2_3D)
    ind = pkg.host.Parms().ind
    dzm = np.where(ref.active_mask(), ref.cell_thickness, 0.0)
    jscale = {}
    for el, slot in (("C", ind.dic_ind), ("N", ind.no3_ind), ("P", ind.po4_ind), ("Si", ind.sio3_ind)):
        # see parity.compare_bgc_source_sink: the conservation residuals are measured against the
        # terms they cancel
        s = float(np.max(np.sum(np.abs(ref.BGC_tendencies[:, :, slot - 1]) * dzm, axis=0)))
        jscale["diag_Jint_%stot" % el] = s
        jscale["diag_Jint_100m_%stot" % el] = s
    for nm, a in ref.diag.items():
        st.add("%s:%s" % (tag, nm), got.diag[nm], a, scale=jscale.get(nm))
    if surface:
        cm = np.arange(ref.nColumnsMax) < ref.nColumns
        for nm, a in ref.forcing.items():
            if a.shape[0] == ref.nColumnsMax:
                st.add("%s:%s" % (tag, nm), got.forcing[nm][cm], a[cm])
        for nm, a in ref.flux_diag.items():
            st.add("%s:%s" % (tag, nm), got.flux_diag[nm][cm], a[cm])


def _active(c):
    k = np.arange(1, c.nLevelsMax + 1)[:, None]
    kmax = c.number_of_active_levels.copy()
    kmax[c.nColumns:] = 0
    return k <= kmax[None, :]


def add_dms(st, tag, ref, got):
    for n in range(abi.DMS_TRACER_CNT):
        st.add("%s:dms_tend[%d]" % (tag, n + 1), got.DMS_tendencies[:, :, n], ref.DMS_tendencies[:, :, n])
    m = _active(ref)
    for nm, a in ref.diag.items():
        st.add("%s:dms.%s" % (tag, nm), got.diag[nm][m], a[m])
    for nm, a in ref.flux_diag.items():
        st.add("%s:dms.%s" % (tag, nm), got.flux_diag[nm], a)
    st.add("%s:dms.netFlux_" % tag, got.forcing["netFlux"], ref.forcing["netFlux"])


def add_macros(st, tag, ref, got):
    for n in range(abi.MACROS_TRACER_CNT):
        st.add("%s:macros_tend[%d]" % (tag, n + 1), got.MACROS_tendencies[:, :, n], ref.MACROS_tendencies[:, :, n])
    m = _active(ref)
    for nm, a in ref.diag.items():
        st.add("%s:macros.%s" % (tag, nm), got.diag[nm][m], a[m])


# ---------------------------------------------------------------- the driver
class MeshChecker:
    """`units`: list of (first column, number of columns) of the device mesh to compare;
    column0: global index of the mesh's first column (the synthetic generator is keyed by it)."""

    def __init__(self, parms, nL, dev_bgc, dev_dms, dev_mac, units, column0=0, ragged=True, nthreads=None):
        self.parms, self.nL = parms, nL
        self.dev = (dev_bgc, dev_dms, dev_mac)
        self.units, self.column0, self.ragged = units, column0, ragged
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        self.nthreads = int(nthreads or max(1, n))
        self.ph = {}            # unit -> the reference's (PH_PREV_3D, PH_PREV_ALT_CO2_3D) after the cold pass
        self._tls = threading.local()

    def _rp(self):
        if not hasattr(self._tls, "rp"):
            self._tls.rp = rt.RefParms(self.parms)
        return self._tls.rp

    def _inputs(self, c0, n):
        bgc = pkg.BgcColumns(self.nL, n)
        dms = pkg.DmsColumns(self.nL, n) if self.dev[1] is not None else None
        mac = pkg.MacrosColumns(self.nL, n) if self.dev[2] is not None else None
        p = self.parms
        pkg.synth_fill(bgc, dms, mac, bgc_ind=p.ind, dms_ind=p.dms_ind if dms else None,
                       macros_ind=p.macros_ind if mac else None, column0=self.column0 + c0, ragged=self.ragged)
        return bgc, dms, mac

    def _cold_unit(self, st, unit):
        c0, n = unit
        rp = self._rp()
        ref, _, _ = self._inputs(c0, n)
        got = ref.copy()
        rt.BGC_SourceSink(rp, ref, True)
        fetch_bgc(self.dev[0], got, slice(c0, c0 + n))
        add_bgc(st, "cold", ref, got, surface=False)
        self.ph[unit] = (ref.PH_PREV_3D.copy(), ref.PH_PREV_ALT_CO2_3D.copy())
        return int(ref.active_mask().sum())

    def _warm_unit(self, st, unit):
        c0, n = unit
        rp = self._rp()
        ref, dref, mref = self._inputs(c0, n)
        ref.PH_PREV_3D[...], ref.PH_PREV_ALT_CO2_3D[...] = self.ph[unit]
        got = ref.copy()
        rt.BGC_SourceSink(rp, ref, True)
        rt.BGC_SurfaceFluxes(rp, ref)
        sl = slice(c0, c0 + n)
        fetch_bgc(self.dev[0], got, sl)
        add_bgc(st, "warm", ref, got, surface=True)
        if dref is not None:
            dgot = dref.copy()
            rt.DMS_SourceSink(rp, dref); rt.DMS_SurfaceFluxes(rp, dref)
            fetch_dms(self.dev[1], dgot, sl)
            add_dms(st, "warm", dref, dgot)
        if mref is not None:
            mgot = mref.copy()
            rt.MACROS_SourceSink(rp, mref)
            fetch_macros(self.dev[2], mgot, sl)
            add_macros(st, "warm", mref, mgot)
        return int(ref.active_mask().sum())

    def _run(self, fn):
        st = FieldStats()
        with ThreadPoolExecutor(max_workers=self.nthreads) as pool:
            cells = sum(pool.map(lambda u: fn(st, u), self.units))
        return st, cells

    def check_cold(self):
        """after ONE BGC_SourceSink of the device mesh with PH_PREV = 0"""
        return self._run(self._cold_unit)

    def check_warm(self):
        """after a second BGC_SourceSink + BGC_SurfaceFluxes + DMS_* + MACROS_SourceSink of the device mesh"""
        return self._run(self._warm_unit)


def units_of(nC, unit, stride=1):
    """[(c0, n)]: every `stride`-th unit of `unit` columns of a mesh of nC columns (the last one shorter)"""
    out = []
    for i, c0 in enumerate(range(0, nC, unit)):
        if i % stride == 0:
            out.append((c0, min(unit, nC - c0)))
    return out
