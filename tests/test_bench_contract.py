"""bench.py's output contract, checked on the CPU arm (the only arm that runs without a GPU):
exactly one line on stdout, valid JSON, the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, FASTRANK_BENCH_N="60000", FASTRANK_BENCH_Q="1800")  # keep the CPU suite short
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["unit"] == "evals/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert set(d["config"]) == {"workload", "evals_per_step", "l2", "parallelism"}  # same keys as the GPU arm
    assert d["details"]["native_so_loaded"] is False


def test_gpu_arm_refuses_without_a_device():
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box the arm runs; covered by the driver's own bench step
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"], cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode != 0
    assert "no CPU fallback" in (res.stderr + res.stdout)
