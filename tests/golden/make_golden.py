"""Regenerates tests/golden/*.npz by RUNNING THE REFERENCE.

The reference ships no tests, fixtures or golden vectors of its own, and the image has no
Fortran compiler; these fixtures are the outputs of the unmodified reference sources
(/root/reference/*.F90) machine-translated to C by oracle/f90c.py and compiled with gcc
(oracle/_ref/libbgc_ref.so, see oracle/ref_translated.py), on the seeded synthetic inputs of
SURVEY.md 8(d).  They serve two purposes: (1) the hand-written oracle must reproduce them bit
for bit on CPU, here and on the GPU box, where /root/reference does not exist
(tests/test_golden.py), (2) the GPU path is checked against fixed numbers as well as against
the live oracle.  `--from-oracle` regenerates them from the oracle instead (for a machine
without the reference sources); the two give identical files.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import parity  # noqa: E402

pkg = parity.pkg
o = oracle_mod = parity.oracle()
sys.path.insert(0, os.path.join(parity.REPO, "oracle"))
import ref_translated as rt   # noqa: E402


class _Ref:
    """the translated reference behind the oracle module's call signatures"""

    def __init__(self):
        self._rp = {}

    def Parms(self):
        po = oracle_mod.Parms()
        self._rp[id(po)] = rt.RefParms(po)
        self._keep = po
        return po

    def BGC_SourceSink(self, po, cols, alt=True):
        rt.BGC_SourceSink(self._rp[id(po)], cols, alt)

    def BGC_SurfaceFluxes(self, po, cols):
        rt.BGC_SurfaceFluxes(self._rp[id(po)], cols)

    def DMS_SourceSink(self, po, cols):
        rt.DMS_SourceSink(self._rp[id(po)], cols)

    def DMS_SurfaceFluxes(self, po, cols):
        rt.DMS_SurfaceFluxes(self._rp[id(po)], cols)

    def MACROS_SourceSink(self, po, cols):
        rt.MACROS_SourceSink(self._rp[id(po)], cols)

    def co2calc_points(self, pts):
        n = len(pts["temp"])
        out = {k: np.zeros(n) for k in ("ph", "co2star", "dco2star", "pco2surf", "dpco2")}
        rt.RefParms(oracle_mod.Parms())
        for i in range(n):
            r = rt.co2calc_1point(*[float(pts[k][i]) for k in ("depth", "temp", "salt", "dic", "ta", "pt",
                                                               "sit", "phlo", "phhi", "xco2", "atmpres")])
            for k in out:
                out[k][i] = r[k]
        return out


if "--from-oracle" not in sys.argv:
    if not rt.available():
        rt.build()
    o = _Ref()
    SOURCE = "translated reference (oracle/_ref/libbgc_ref.so)"
else:
    SOURCE = "oracle (oracle/libbgc_oracle.so)"


def single_column():
    """BASELINE.json configs[0]: one column, 60 levels, idealised profiles (no jitter)."""
    po = o.Parms()
    cols, dms, mac = parity.make_bgc(60, 1, po, jitter=False, with_dms=True, with_macros=True)
    o.BGC_SourceSink(po, cols, True)
    cold_ph = cols.PH_PREV_3D.copy()
    o.BGC_SourceSink(po, cols, True)     # warm-bracket pass
    o.BGC_SurfaceFluxes(po, cols)
    o.DMS_SourceSink(po, dms); o.DMS_SurfaceFluxes(po, dms)
    o.MACROS_SourceSink(po, mac)
    out = {"tend": cols.BGC_tendencies[:, 0, :], "ph_cold": cold_ph[:, 0], "ph_warm": cols.PH_PREV_3D[:, 0],
           "netFlux": cols.forcing["netFlux"][0], "surface_pH": cols.forcing["surface_pH"],
           "dms_tend": dms.DMS_tendencies[:, 0, :], "dms_netFlux": dms.forcing["netFlux"][0],
           "macros_tend": mac.MACROS_tendencies[:, 0, :]}
    for nm in ("diag_CO3", "diag_co3_sat_calc", "diag_PAR_avg", "diag_POC_FLUX_IN", "diag_POC_REMIN",
               "diag_P_iron_REMIN", "diag_NITRIF", "diag_DENITRIF", "diag_O2_CONSUMPTION", "diag_AOU",
               "diag_SedDenitrif", "diag_pocToSed"):
        out[nm] = cols.diag[nm][:, 0]
    for nm in ("diag_photoC", "diag_light_lim", "diag_auto_graze"):
        out[nm] = cols.diag[nm][:, 0, :]
    for nm in pkg.abi.BGC_DIAG_C1:
        out[nm] = cols.diag[nm]
    np.savez_compressed(os.path.join(HERE, "single_column_60.npz"), **out)


def co2_points():
    """BASELINE.json configs[1] (first 512 of the 1M points)."""
    pts = pkg.synth_co2_points(512)
    r = o.co2calc_points(pts)
    np.savez_compressed(os.path.join(HERE, "co2calc_points_512.npz"),
                        **{k: r[k] for k in ("ph", "co2star", "dco2star", "pco2surf", "dpco2")})


def ragged_block():
    """A 24-level x 96-column ragged block with jitter (branch coverage)."""
    po = o.Parms()
    cols, _, _ = parity.make_bgc(24, 96, po, ragged=True, nColumns=90)
    o.BGC_SourceSink(po, cols, True)
    np.savez_compressed(os.path.join(HERE, "ragged_block_24x96.npz"),
                        tend=cols.BGC_tendencies, ph=cols.PH_PREV_3D,
                        jint=np.stack([cols.diag["diag_Jint_100m_%stot" % e] for e in ("C", "N", "P", "Si")]),
                        zsat=np.stack([cols.diag["diag_zsatcalc"], cols.diag["diag_zsatarag"]]),
                        o2min=np.stack([cols.diag["diag_O2_ZMIN"], cols.diag["diag_O2_ZMIN_DEPTH"]]))


def surface_dms_macros_block():
    """Everything except BGC_SourceSink on a ragged 20 x 64 block: BGC_SurfaceFluxes with its in-place
    side effects on the forcing, DMS_SourceSink / DMS_SurfaceFluxes, MACROS_SourceSink (their
    diagnostics on active cells only: the reference leaves the rest untouched)."""
    po = o.Parms()
    cols, dms, mac = parity.make_bgc(20, 64, po, ragged=True, nColumns=61, with_dms=True, with_macros=True)
    cols.forcing["iceFraction"][:7] = [-0.2, 1.4, 0.3, 0.0, 1.0, 2.0, -1.0]
    o.BGC_SurfaceFluxes(po, cols)
    o.DMS_SourceSink(po, dms); o.DMS_SurfaceFluxes(po, dms)
    o.MACROS_SourceSink(po, mac)
    act = dms.active_mask() if hasattr(dms, "active_mask") else cols.active_mask()
    out = {"netFlux": cols.forcing["netFlux"], "gasFlux": cols.forcing["gasFlux"],
           "iceFraction": cols.forcing["iceFraction"], "surface_pH": cols.forcing["surface_pH"],
           "depositionFlux": cols.forcing["depositionFlux"],
           "dms_tend": dms.DMS_tendencies, "dms_netFlux": dms.forcing["netFlux"],
           "macros_tend": mac.MACROS_tendencies}
    for n, a in cols.flux_diag.items():
        out["bflux_" + n] = a
    for n, a in dms.diag.items():
        out["dms_" + n] = np.where(act, a, 0.0)
    for n, a in dms.flux_diag.items():
        out["dflux_" + n] = a[:61]
    for n, a in mac.diag.items():
        out["mac_" + n] = np.where(act, a, 0.0)
    np.savez_compressed(os.path.join(HERE, "surface_dms_macros_20x64.npz"), **out)


if __name__ == "__main__":
    single_column(); co2_points(); ragged_block(); surface_dms_macros_block()
    print("golden fixtures written to", HERE, "from the", SOURCE)
