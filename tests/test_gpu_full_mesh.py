"""The BASELINE.json configurations AT THEIR FULL SIZE, cell for cell against the reference.

  * config 4 (the headline): the whole EC60to30 mesh, 235 160 columns x 60 levels, device resident,
    ONE C-ABI call per procedure - BGC_SourceSink cold, then the full warm step
    (BGC_SourceSink + BGC_SurfaceFluxes + DMS_SourceSink + DMS_SurfaceFluxes + MACROS_SourceSink) -
    every tendency, both pH fields, every diagnostic and every forcing side effect compared with
    the translated reference (oracle/_ref/libbgc_ref.so) run on all host threads, at the
    tolerances of tests/parity.py;
  * config 5: one GPU's slab of the RRS18to6 mesh (461 654 columns x 80 levels, ~82 GB resident), the
    same comparison on every 16th unit of 512 columns;
  * config 2: co2calc_1point over the full 1 048 576 points against the OpenMP oracle (the oracle
    equals the translated reference bit for bit on this procedure: tests/test_reference_translated.py);
  * a 20-step trajectory with evolving tracers (T += dt * tendency), GPU against the reference.

CPU part (-m "not gpu"): the comparison harness itself, with the oracle standing in for the GPU.
"""
import os
import sys
import time

import numpy as np
import pytest

import parity
import mesh_parity as mp

pkg = parity.pkg
abi = pkg.abi
rt = mp.rt

need_ref = pytest.mark.skipif(not rt.available(), reason="oracle/_ref/libbgc_ref.so not shipped")


def _host():
    return pkg.host


def _nthreads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------ CPU: the harness itself
@need_ref
def test_mesh_checker_with_the_oracle_as_the_device():
    """Dry run of tests/mesh_parity.py: the "device" containers are CPU tensors and the oracle plays
    the GPU.  Checks synth_fill_device (SoA generator, ragged) against synth_fill (Fortran layout),
    the unit-wise fetch / compare plumbing, and that a planted error is caught."""
    host = _host()
    o = parity.oracle()
    po = o.Parms()
    parms = host.Parms()
    nL, nC = 18, 300
    dev = host.DeviceBgcColumns(nL, nC, device="cpu")
    ddev = host.DeviceDmsColumns(nL, nC, device="cpu")
    mdev = host.DeviceMacrosColumns(nL, nC, device="cpu")
    cells = pkg.synth_fill_device(parms, dev, ddev, mdev, column0=1000, ragged=True)
    h, hd, hm = pkg.BgcColumns(nL, nC), pkg.DmsColumns(nL, nC), pkg.MacrosColumns(nL, nC)
    pkg.synth_fill(h, hd, hm, bgc_ind=parms.ind, dms_ind=parms.dms_ind, macros_ind=parms.macros_ind,
                   column0=1000, ragged=True)
    assert cells == int(h.active_mask().sum())
    assert np.array_equal(dev.BGC_tracers.numpy(), np.transpose(h.BGC_tracers, (2, 0, 1)))
    assert np.array_equal(dev.cell_thickness.numpy(), h.cell_thickness)
    assert np.array_equal(dev.number_of_active_levels.numpy(), h.number_of_active_levels)
    assert np.array_equal(ddev.DMS_tracers.numpy(), np.transpose(hd.DMS_tracers, (2, 0, 1)))
    assert np.array_equal(mdev.MACROS_tracers.numpy(), np.transpose(hm.MACROS_tracers, (2, 0, 1)))
    for n, t in dev.forcing.items():
        a = h.forcing[n]
        assert np.array_equal(t.numpy(), a.T if n in abi.BGC_FORCING_FLUX else a), n

    units = mp.units_of(nC, 64)
    assert units[-1] == (256, 44)
    chk = mp.MeshChecker(parms, nL, dev, ddev, mdev, units, column0=1000, ragged=True, nthreads=4)
    o.BGC_SourceSink(po, h, True)
    dev.load(h)
    st, n = chk.check_cold()
    assert n == cells
    errs = st.check("cold pass")
    assert max(errs.values()) == 0.0          # oracle == translated reference, bit for bit
    o.BGC_SourceSink(po, h, True); o.BGC_SurfaceFluxes(po, h)
    o.DMS_SourceSink(po, hd); o.DMS_SurfaceFluxes(po, hd); o.MACROS_SourceSink(po, hm)
    dev.load(h); ddev.load(hd); mdev.load(hm)
    st, _ = chk.check_warm()
    errs = st.check("warm step")
    assert max(errs.values()) == 0.0 and len(errs) > 200
    # a planted error in one cell of one array is found
    dev.diag["diag_POC_REMIN"][5, 290] += 1e-3
    with pytest.raises(AssertionError, match="diag_POC_REMIN"):
        chk.check_warm()[0].check("planted")


# ------------------------------------------------------------------ GPU
def _device_step(host, ctx, bgc, dms, mac):
    host.BGC_SourceSink(ctx, bgc, True, True)
    host.BGC_SurfaceFluxes(ctx, bgc)
    host.DMS_SourceSink(ctx, dms, True)
    host.DMS_SurfaceFluxes(ctx, dms)
    host.MACROS_SourceSink(ctx, mac, True)
    ctx.synchronize()


def _full_mesh(nL, nC, unit, stride, column0, what):
    host = _host()
    parms = host.Parms()
    ctx = host.Context(nL, nC, device=0, parms=parms)
    bgc = host.DeviceBgcColumns(nL, nC)
    dms = host.DeviceDmsColumns(nL, nC)
    mac = host.DeviceMacrosColumns(nL, nC)
    t0 = time.time()
    cells = pkg.synth_fill_device(parms, bgc, dms, mac, column0=column0, ragged=True, nthreads=_nthreads())
    t_fill = time.time() - t0
    units = mp.units_of(nC, unit, stride)
    chk = mp.MeshChecker(parms, nL, bgc, dms, mac, units, column0=column0, ragged=True)
    host.BGC_SourceSink(ctx, bgc, True, True)      # cold brackets: PH_PREV = 0
    ctx.synchronize()
    t0 = time.time()
    st, n_cold = chk.check_cold()
    t_cold = time.time() - t0
    e_cold = st.check(what + ", cold pass")
    _device_step(host, ctx, bgc, dms, mac)         # warm brackets, the whole step
    t0 = time.time()
    st, n_warm = chk.check_warm()
    t_warm = time.time() - t0
    e_warm = st.check(what + ", warm step")
    status = ctx.status()
    assert status["no_bracket"] == 0 and status["no_convergence"] == 0 and status["nonfinite"] == 0, status
    ctx.close()
    line = ("%s: %d active cells on the device, %d compared (cold) / %d (warm) in %d units; fill %.1f s, "
            "reference cold %.1f s, warm %.1f s on %d threads; %d + %d arrays; worst error cold %.2e (%s), warm %.2e (%s)"
            % (what, cells, n_cold, n_warm, len(units), t_fill, t_cold, t_warm, chk.nthreads, len(e_cold), len(e_warm),
               max(e_cold.values()), max(e_cold, key=e_cold.get), max(e_warm.values()), max(e_warm, key=e_warm.get)))
    print("\n" + line)
    out = os.path.join(parity.REPO, "gpurun_out")
    if os.path.isdir(out):     # the box's scratch directory: comes back with the run
        with open(os.path.join(out, "full_mesh_parity.txt"), "a") as f:
            f.write(line + "\n")
    return cells, n_cold, e_cold, e_warm


@pytest.mark.gpu
@need_ref
def test_ec60to30_full_mesh_against_the_reference():
    """BASELINE.json configs[3] = the bench workload, every cell (BGC_mod.F90:340-1998 & co)."""
    cells, n, e_cold, e_warm = _full_mesh(60, 235160, 1024, 1, 0, "EC60to30 235160 x 60")
    assert n == cells                         # every active cell of the mesh was compared
    assert len(e_warm) > 200


@pytest.mark.gpu
@need_ref
def test_rrs18to6_slab_against_the_reference():
    """BASELINE.json configs[4]: rank 3's slab of the 8-GPU split of RRS18to6 (461 654 columns x 80
    levels in one call, ~82 GB resident); every 16th unit of 512 columns is compared."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 100e9:
        pytest.skip("needs ~90 GB of free device memory")
    nC = -(-3693225 // 8)
    cells, n, e_cold, e_warm = _full_mesh(80, nC, 512, 16, 3 * nC, "RRS18to6 slab %d x 80" % nC)
    assert n > cells // 20


@pytest.mark.gpu
def test_co2calc_one_million_points():
    """BASELINE.json configs[1] at its full size (co2calc.F90:75-210), cold and warm brackets."""
    host = _host()
    o = parity.oracle()
    parms = host.Parms()
    ctx = host.Context(2, 64, device=0, parms=parms)
    n = 1 << 20
    pts = pkg.synth_co2_points(n)
    for warm in (False, True):
        r = o.co2calc_points(pts, nthreads=o.max_threads())
        g = host.co2calc_points(ctx, pts)
        for k in ("ph", "co2star", "dco2star", "pco2surf", "dpco2"):
            e = parity.nerr(g[k], r[k])
            assert e <= parity.TOL_SOLVER, (warm, k, e)
        dH = np.max(np.abs(10.0 ** -g["ph"] - 10.0 ** -r["ph"]))
        assert dH <= 1e-10, dH                # the solver's own stopping tolerance, mol/kg
        pts["phlo"], pts["phhi"] = r["ph"] - 0.2, r["ph"] + 0.2
    st = ctx.status()
    assert st["no_bracket"] == 0 and st["no_convergence"] == 0, st
    ctx.close()


@pytest.mark.gpu
@need_ref
@pytest.mark.parametrize("device_mode", [True, False])
def test_twenty_step_trajectory(device_mode):
    """GPU and reference each advance their OWN tracers for 20 steps (T += dt * tendency, PH_PREV
    carried from step to step, surface pH fed by BGC_SurfaceFluxes): the solver-dependent state that
    is held to 1e-8 per step must not drift.  After 20 steps the tracers agree to 1e-9 of their
    magnitude and the last step's tendencies to 1e-8."""
    host = _host()
    parms = host.Parms()
    rp = rt.RefParms(parms)
    nL, nC, dt = 60, 384, 1800.0
    ctx = host.Context(nL, nC, device=0, parms=parms)
    cols, _, _ = parity.make_bgc(nL, nC, parms, ragged=True)
    ref, got = cols.copy(), cols.copy()
    dev = host.DeviceBgcColumns(nL, nC).load(got) if device_mode else None
    active = ref.active_mask()[:, :, None]
    for step in range(20):
        rt.BGC_SourceSink(rp, ref, True)
        rt.BGC_SurfaceFluxes(rp, ref)
        ref.BGC_tracers[...] = np.where(active, ref.BGC_tracers + dt * ref.BGC_tendencies, ref.BGC_tracers)
        if device_mode:
            host.BGC_SourceSink(ctx, dev, True, True)
            host.BGC_SurfaceFluxes(ctx, dev)
            ctx.synchronize()
            am = dev.BGC_tracers.new_tensor(np.transpose(active, (2, 0, 1)).astype(np.float64))
            dev.BGC_tracers += dt * dev.BGC_tendencies * am
            import torch
            torch.cuda.synchronize()
        else:
            host.BGC_SourceSink(ctx, got, True, True)
            host.BGC_SurfaceFluxes(ctx, got)
            got.BGC_tracers[...] = np.where(active, got.BGC_tracers + dt * got.BGC_tendencies, got.BGC_tracers)
    if device_mode:
        dev.store(got)
        got.BGC_tracers[...] = np.transpose(dev.BGC_tracers.cpu().numpy(), (1, 2, 0))
    for n in range(abi.BGC_TRACER_CNT):
        e = parity.nerr(got.BGC_tracers[:, :, n], ref.BGC_tracers[:, :, n])
        assert e <= 1e-9, ("tracer", n + 1, e)
    parity.compare_bgc_source_sink(ref, got, tol=1e-8, tol_solver=1e-7)
    dH = np.max(np.abs(np.where(active[:, :, 0], 10.0 ** -got.PH_PREV_3D - 10.0 ** -ref.PH_PREV_3D, 0.0)))
    assert dH <= 2e-10, dH
    st = ctx.status()
    assert st["no_bracket"] == 0 and st["no_convergence"] == 0 and st["nonfinite"] == 0, st
    ctx.close()
