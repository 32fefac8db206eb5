"""Step time against mesh size for the three placements of the carbonate kernel
(bgc_ctx_set_concurrency: 0 = on the ctx stream after the sweep, 1 = side stream, placement by sweep
size (the default), 2 = side stream forked after the sweep, beside the DMS / MACROS / surface kernels,
3 = confined to the SMs a sub-wave sweep leaves idle; BGC_CO3_CONFINED=0 turns the confinement off in
mode 1, BGC_CO3_SHARE=<percent> scales the share of the cells given to the confined blocks).
    python scripts/micro/concurrency_sweep.py > gpurun_out/concurrency_sweep.txt"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import torch  # noqa: E402

pkg = bench.ge.load_package()
parms = pkg.host.Parms()
sizes = [int(x) for x in sys.argv[1:]] or [29396, 58790, 117580, 235160]
for nC in sizes:
    mesh = bench.DeviceMesh(pkg, parms, 60, nC, 0, 0, 0, 1)
    for _ in range(5):
        mesh.step()
    row = []
    for mode in (0, 1, 2, 3):
        mesh.ctx.set_concurrency(mode)
        for _ in range(3):
            mesh.step()
        g = mesh.capture()
        ms = min(mesh.timed(50, g) for _ in range(3))
        mesh.ctx.graph_destroy(g)
        row.append(ms)
    print("columns %7d  same-stream %.4f  default %.4f  after-sweep %.4f  confined %.4f ms/step  (share %s, confined-by-default %s)"
          % (nC, *row, os.environ.get("BGC_CO3_SHARE", "85"), os.environ.get("BGC_CO3_CONFINED", "1")), flush=True)
    mesh.close()
    del mesh
    torch.cuda.empty_cache()
