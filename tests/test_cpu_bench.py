"""bench.py's CPU-runnable arms print exactly one JSON line carrying the contract's keys
(the CUDA arm needs a GPU and is exercised by the driver / profiles/)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
        "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline",
        "e2e", "gpu_launches"}  # fmt: skip


@pytest.mark.parametrize("extra", [["--cpu-sample-chains", "256"],
                                   ["--workload", "sv", "--cpu-sample-chains", "1"]])  # fmt: skip
def test_reference_arm_prints_one_json_line(extra):
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
         "--warmup", "0", *extra], capture_output=True, text=True, timeout=600, cwd=ROOT)  # fmt: skip
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d) and d["impl"] == "reference" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["dtype"] == "f64"


def test_cuda_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)  # fmt: skip
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_under_two_ranks_prints_once():
    """The driver launches `--impl reference` like the CUDA arm (torchrun, one process per
    GPU): rank 0 alone runs and prints the line - `n_gpus` = 2, the strong-scaling config of
    the 2-GPU run -, the other rank exits 0 without work."""
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "bench.py"),
         "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
         "--cpu-sample-chains", "256"], capture_output=True, text=True, timeout=900, cwd=ROOT)  # fmt: skip
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d) and d["impl"] == "reference" and d["n_gpus"] == 2
    assert d["config"]["chains_total"] == 65536 and d["config"]["chains_per_gpu"] == 32768
