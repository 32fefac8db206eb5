"""Poor man's timeline of one step: every launch of the library bracketed by CUDA events on its own
stream (BGC_TRACE_FILE), with the carbonate side stream running concurrently as in production.
    python scripts/micro/trace_step.py [columns] > gpurun_out/trace.txt"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
out = os.path.join(REPO, "gpurun_out", "trace_raw.txt")
if os.path.exists(out):
    os.remove(out)
os.environ["BGC_TRACE_FILE"] = out
import bench  # noqa: E402
import torch  # noqa: E402

nC = int(sys.argv[1]) if len(sys.argv) > 1 else 29396
pkg = bench.ge.load_package()
parms = pkg.host.Parms()
mesh = bench.DeviceMesh(pkg, parms, 60, nC, 0, 0, 0, 1)
for _ in range(5):
    mesh.step()
mesh.ctx.synchronize()
mesh.ctx.timing_reset()
mesh.ctx.timing_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(mesh.stream)
for _ in range(3):
    mesh.step()
e1.record(mesh.stream)
mesh.ctx.synchronize()
mesh.ctx.timing()
print("3 steps with per-launch events: %.4f ms per step" % (e0.elapsed_time(e1) / 3))
print(open(out).read())
mesh.ctx.timing_enable(False)
ms = mesh.timed(100)
print("eager, no events: %.4f ms per step" % ms)
g = mesh.capture()
print("graph: %.4f ms per step" % mesh.timed(100, g))
