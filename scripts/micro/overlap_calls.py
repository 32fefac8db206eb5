"""Would running DMS_SourceSink / MACROS_SourceSink BESIDE BGC_SourceSink pay when a GPU owns less than
one wave of sweep blocks?  Emulated with two ctxs (two streams) on one device.
    python scripts/micro/overlap_calls.py [columns ...]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import torch  # noqa: E402

pkg = bench.ge.load_package()
host = pkg.host
parms = host.Parms()
for nC in [int(x) for x in sys.argv[1:]] or [29396, 58790, 235160]:
    mesh = bench.DeviceMesh(pkg, parms, 60, nC, 0, 0, 0, 1, inventory=False)
    ctx2 = host.Context(60, nC, device=0, parms=parms)
    s2 = torch.cuda.Stream()
    ctx2.set_stream(s2.cuda_stream)

    def serial():
        mesh.compute()

    def overlapped(order):
        ev = torch.cuda.Event()
        ev.record(mesh.stream)
        s2.wait_event(ev)
        if order == "dms_first":
            host.DMS_SourceSink(ctx2, mesh.dms, True); host.DMS_SurfaceFluxes(ctx2, mesh.dms)
            host.MACROS_SourceSink(ctx2, mesh.mac, True)
        host.BGC_SourceSink(mesh.ctx, mesh.bgc, True, True)
        host.BGC_SurfaceFluxes(mesh.ctx, mesh.bgc)
        if order != "dms_first":
            host.DMS_SourceSink(ctx2, mesh.dms, True); host.DMS_SurfaceFluxes(ctx2, mesh.dms)
            host.MACROS_SourceSink(ctx2, mesh.mac, True)
        mesh.ctx.carbonate_join()
        ev2 = torch.cuda.Event()
        ev2.record(s2)
        mesh.stream.wait_event(ev2)

    def timed(fn, reps=100):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(mesh.stream)
        for _ in range(reps):
            fn()
        e1.record(mesh.stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    for _ in range(3):
        serial(); overlapped("bgc_first")
    print("columns %7d  serial %.4f  DMS+MACROS beside the sweep (BGC issued first) %.4f  (DMS issued first) %.4f ms/step"
          % (nC, timed(serial), timed(lambda: overlapped("bgc_first")), timed(lambda: overlapped("dms_first"))), flush=True)
    ctx2.close()
    mesh.close()
    del mesh
    torch.cuda.empty_cache()
