"""One-off, larger comparison of the hand-written oracle with the translated reference
(oracle/_ref/libbgc_ref.so): three slabs of the synthetic meshes (2.4 M cells, cold + warm pass,
surface fluxes, DMS, MACROS; every tendency, pH field, diagnostic and forcing side effect) and
100 000 config-2 points.  Expected output: "bit-identical: True" on every line, 0 mismatches.
Run from the repository root:  python scripts/compare_oracle_reference.py   (about 20 s)
"""
import sys, time
sys.path.insert(0,'tests'); sys.path.insert(0,'oracle')
import parity, numpy as np
o = parity.oracle(); import ref_translated as rt
po = o.Parms(); rp = rt.RefParms(po)
pkg = parity.pkg
tot = 0
for (nL, nC, ragged, c0) in [(60, 16384, False, 0), (60, 16384, True, 100000), (80, 8192, True, 3000000)]:
    bgc, dms, mac = pkg.BgcColumns(nL, nC), pkg.DmsColumns(nL, nC), pkg.MacrosColumns(nL, nC)
    pkg.synth_fill(bgc, dms, mac, bgc_ind=po.ind, dms_ind=po.dms_ind, macros_ind=po.macros_ind, column0=c0, ragged=ragged)
    a, b = bgc.copy(), bgc.copy(); da, db = dms.copy(), dms.copy(); ma, mb = mac.copy(), mac.copy()
    for p in range(2):
        o.BGC_SourceSink(po, a, True, nthreads=8); rt.BGC_SourceSink(rp, b, True)
    o.BGC_SurfaceFluxes(po, a, nthreads=8); rt.BGC_SurfaceFluxes(rp, b)
    o.DMS_SourceSink(po, da, nthreads=8); rt.DMS_SourceSink(rp, db); o.DMS_SurfaceFluxes(po, da); rt.DMS_SurfaceFluxes(rp, db)
    o.MACROS_SourceSink(po, ma, nthreads=8); rt.MACROS_SourceSink(rp, mb)
    ok = np.array_equal(a.BGC_tendencies, b.BGC_tendencies) and np.array_equal(a.PH_PREV_3D, b.PH_PREV_3D)
    ok &= all(np.array_equal(a.diag[n], b.diag[n]) for n in a.diag)
    ok &= all(np.array_equal(a.forcing[n], b.forcing[n]) for n in a.forcing) and all(np.array_equal(a.flux_diag[n], b.flux_diag[n]) for n in a.flux_diag)
    ok &= np.array_equal(da.DMS_tendencies, db.DMS_tendencies) and all(np.array_equal(da.diag[n], db.diag[n]) for n in da.diag)
    ok &= np.array_equal(ma.MACROS_tendencies, mb.MACROS_tendencies) and all(np.array_equal(ma.diag[n], mb.diag[n]) for n in ma.diag)
    cells = int(bgc.active_mask().sum()); tot += cells
    print(nL, nC, ragged, c0, "cells", cells, "bit-identical:", ok, flush=True)
print("total cells", tot)
pts = pkg.synth_co2_points(100000)
g = o.co2calc_points(pts, nthreads=8)
t0 = time.time()
bad = 0
for i in range(100000):
    r = rt.co2calc_1point(*[float(pts[k][i]) for k in ("depth","temp","salt","dic","ta","pt","sit","phlo","phhi","xco2","atmpres")])
    for k in ("ph","co2star","dco2star","pco2surf","dpco2"):
        if r[k] != g[k][i]: bad += 1
print("co2calc 100000 points, mismatching values:", bad, "%.1fs" % (time.time()-t0))
