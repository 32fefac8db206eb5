"""Regenerates tests/golden/*.npz from the CPU oracle.

The reference ships no tests, fixtures or golden vectors and cannot be built in
this image (no Fortran compiler), so these fixtures do NOT come from the
reference itself: they freeze the oracle's answers at the commit that produced
them.  They serve two purposes: (1) any later edit of the oracle that changes
its results is caught on CPU (tests/test_golden.py), (2) the GPU path is checked
against fixed numbers on the GPU box as well as against the live oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import parity  # noqa: E402

pkg = parity.pkg
o = parity.oracle()


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


if __name__ == "__main__":
    single_column(); co2_points(); ragged_block()
    print("golden fixtures written to", HERE)
