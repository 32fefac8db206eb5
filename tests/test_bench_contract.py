"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the
agreed keys, and the B200 arm refuses to run (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], capture_output=True, text=True,
                          cwd=REPO, env=e, timeout=600)


def test_reference_arm_prints_one_json_line(built):
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-columns", "256", "--gpus", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "cell-updates/s" and d["dtype"] == "f64"
    assert d["vs_baseline"] is None and d["higher_is_better"] is True
    # "reference" when oracle/_ref holds the translated reference (built wherever the reference
    # sources exist and shipped with the snapshot), otherwise the oracle port
    ref_built = os.path.exists(os.path.join(REPO, "oracle", "_ref", "libbgc_ref.so"))
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_built else "port")
    assert d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and "workload" in d["config"]


def test_reference_arm_can_time_the_port(built):
    """BGC_BENCH_CPU_KIND=port (or a missing oracle/_ref) times the hand-written oracle and says so"""
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-columns", "256", "--gpus", "1",
             env={"BGC_BENCH_CPU_KIND": "port"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_reference_arm_other_ranks_exit_quietly(built):
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-columns", "256", "--gpus", "2",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


def test_traffic_json_is_what_its_ncu_summary_says(tmp_path):
    """bench.py quotes roofline.traffic / ncu_counters from profiles/traffic.json (ncu cannot run inside a
    timed bench).  The file names the ncu summary it was made from; regenerating it from that summary
    (scripts/make_traffic_json.py) must give the same numbers, and the summary must be committed."""
    tj = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))
    src = tj["source"].split(" ")[0]
    assert os.path.exists(os.path.join(REPO, src)), src
    out = tmp_path / "traffic.json"
    subprocess.run([sys.executable, os.path.join(REPO, "scripts", "make_traffic_json.py"), src, tj["commit"], str(out)],
                   cwd=REPO, check=True, capture_output=True)
    assert json.load(open(out)) == tj
    # the sweep's measured DRAM traffic is its algorithmic bytes (1464 B per cell): no wasted re-reads
    assert 1464 <= tj["eco_columns_kernel_bytes_per_cell"] <= 1464 * 1.01
