"""Where and when the thread blocks of the column sweep (kernel 1) and of the confined carbonate blocks
(kernel 2) ran: BGC_BLOCK_TRACE_FILE (bgc_kernels.cuh: block_trace_begin).  Prints, for the LAST traced
step, per kernel the number of blocks, the SMs used, and the spread of start times and durations.
    python scripts/micro/block_trace.py [columns] [steps]"""
import os
import sys
import collections

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
out = os.path.join(REPO, "gpurun_out", "block_trace_raw.txt")
os.makedirs(os.path.dirname(out), exist_ok=True)
if os.path.exists(out):
    os.remove(out)
os.environ["BGC_BLOCK_TRACE_FILE"] = out
import bench  # noqa: E402

nC = int(sys.argv[1]) if len(sys.argv) > 1 else 29396
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
pkg = bench.ge.load_package()
parms = pkg.host.Parms()
mesh = bench.DeviceMesh(pkg, parms, 60, nC, 0, 0, 0, 1)
for _ in range(steps):
    mesh.step()
mesh.ctx.synchronize()
mesh.close()
rows = [tuple(int(x) for x in ln.split()) for ln in open(out)]
# split into launches: a new launch of kernel 1 starts where blockIdx 0 of kernel 1 appears
sweeps = [i for i, r in enumerate(rows) if r[0] == 1 and r[4] == 0]
last = rows[sweeps[-1] - (sum(1 for r in rows if r[0] == 2) // max(1, len(sweeps))):]
t0 = min(r[2] for r in last)
for kid, name in ((1, "sweep"), (2, "confined carbonate")):
    rs = [r for r in last if r[0] == kid]
    if not rs:
        continue
    sms = collections.Counter(r[1] for r in rs)
    st = sorted((r[2] - t0) / 1e3 for r in rs)
    du = sorted((r[3] - r[2]) / 1e3 for r in rs)
    en = max((r[3] - t0) / 1e3 for r in rs)
    print("%-20s blocks %4d on %3d SMs; start us min/median/max %.1f %.1f %.1f; duration us min/median/max %.1f %.1f %.1f; last end %.1f"
          % (name, len(rs), len(sms), st[0], st[len(st) // 2], st[-1], du[0], du[len(du) // 2], du[-1], en))
both = collections.defaultdict(set)
for r in last:
    both[r[1]].add(r[0])
print("SMs that ran both kernels in this step:", sum(1 for v in both.values() if len(v) == 2))
sw = sorted((r for r in last if r[0] == 1), key=lambda r: r[3] - r[2])
co = {r[1] for r in last if r[0] == 2}
print("slowest sweep blocks (SM, start us, duration us, TPC sibling runs carbonate):")
for r in sw[-8:]:
    print("   SM %3d  %.1f  %.1f  %s" % (r[1], (r[2] - t0) / 1e3, (r[3] - r[2]) / 1e3, (r[1] ^ 1) in co))
print("fastest:")
for r in sw[:4]:
    print("   SM %3d  %.1f  %.1f  %s" % (r[1], (r[2] - t0) / 1e3, (r[3] - r[2]) / 1e3, (r[1] ^ 1) in co))
# duration against "distance" to carbonate SMs
by = collections.defaultdict(list)
for r in sw:
    by[(r[1] ^ 1) in co].append((r[3] - r[2]) / 1e3)
for k, v in by.items():
    print("sweep blocks whose TPC sibling %s carbonate: %d, mean duration %.1f us" % ("runs" if k else "does not run", len(v), sum(v) / len(v)))
