"""Times the MPAS-layout adapter kernels (bench.run_mpas_adapter) for a list of tuning settings, one
subprocess each (the library reads BGC_MPAS_* once).  Usage on a GPU box:
    python scripts/micro/mpas_sweep.py > gpurun_out/mpas_sweep.txt"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PROG = r'''
import sys, json
sys.path.insert(0, %r)
import bench, torch
pkg = bench.ge.load_package(); host = pkg.host
nL, nC = 60, 235160
ctx = host.Context(nL, nC, device=0, parms=host.Parms())
st = torch.cuda.Stream(); ctx.set_stream(st.cuda_stream)
print(json.dumps(bench.run_mpas_adapter(ctx, st, "cuda:0", nL, nC)))
''' % REPO

SETTINGS = [{"BGC_MPAS_VARIANT": "0"}, {"BGC_MPAS_VARIANT": "1"}] + [
    {"BGC_MPAS_VARIANT": "0", "BGC_MPAS_KB": str(kb), "BGC_MPAS_BLOCKS_PER_SM": str(b)}
    for kb, b in ((2, 3), (2, 4), (3, 2), (3, 3), (4, 2), (4, 3), (6, 2))]
for s in SETTINGS:
    env = dict(os.environ, **s)
    r = subprocess.run([sys.executable, "-c", PROG], capture_output=True, text=True, env=env)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(s, {k: (round(v["ms"], 3), round(v["GBps"])) for k, v in d.items()}, flush=True)
    except Exception:
        print(s, "FAILED", r.stderr[-400:], flush=True)
