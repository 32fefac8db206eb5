"""N > 1 on hardware: the all-reduced 64-double inventory of a mesh split over the GPUs of the box
against the vector ONE GPU computes for the whole mesh (tests/multi_gpu_worker.py under torchrun).
Skipped on a box with a single GPU (the driver's GPU test box); run with `gpurun --gpus 2`, log under
profiles/.  The CPU counterpart (gloo, world size 2) is tests/test_multi_gpu_gloo.py."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpus() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("even", [1, 0])
def test_allreduced_inventory_equals_the_single_gpu_vector(even):
    world = min(_gpus(), 8)
    env = dict(os.environ, BGC_TEST_MESH_COLUMNS="6001", BGC_TEST_EVEN=str(even))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(29711 + even),
                        os.path.join(REPO, "tests", "multi_gpu_worker.py")],
                       capture_output=True, text=True, env=env, timeout=900, cwd=REPO)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    print(d)
    assert d["world"] == world
    assert d["counts"] == d["counts_single"]                  # active cells and columns: exact
    assert d["zero_pattern_equal"] and d["nonzero_sums"] >= 35
    # per-column results are identical whatever the sharding: only the order of the additions differs
    # (sums of ~1e5 terms of mixed sign: 1e-10 relative is many orders above what reordering does)
    assert d["rel_vs_single_gpu"] <= 1e-10, d
    assert d["jint_abs_diff"] <= 1e-10 * d["jint_scale"], d
    assert d["rel_vs_gathered"] <= 1e-12, d
    assert d["rel_graph_replay_vs_eager"] <= 1e-12, d
