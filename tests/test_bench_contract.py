"""bench.py's contract that can be checked without a GPU: the reference arm's JSON line (it is the oracle on the host
cores, so it runs here), the loud failure of our arm when there is no CUDA device, and the multi-GPU layout flags."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*argv, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


@pytest.mark.timeout(900)
def test_reference_arm_line(oracle):
    p = _run("--impl", "reference", "--config", "c1", "--steps", "2", "--warmup", "1")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert d["metric"].startswith("QPS at recall@10 parity")
    assert "c1: 1M x 128, IVF1024,PQ16x8, nprobe=16, k=10" in d["config"]["workload"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["config"]["index"] == "IVF1024,PQ16x8" and d["config"]["batch"] == 10000
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.timeout(900)
def test_reference_arm_uses_all_cores_under_torchrun(oracle):
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arm must still use every host core (round 1's N >= 2
    ratios were inflated by a single-threaded reference)."""
    p = _run("--impl", "reference", "--config", "c1", "--steps", "1", "--warmup", "1", "--nq", "200",
             env={"OMP_NUM_THREADS": "1"})
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads([ln for ln in p.stdout.splitlines() if ln.strip()][0])
    ncpu = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == ncpu, (d["cpu_baseline"], ncpu)


def test_our_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = _run("--steps", "1", "--warmup", "1")
    assert p.returncode != 0
    assert "CUDA" in (p.stderr + p.stdout)
    assert not [ln for ln in p.stdout.splitlines() if ln.startswith("{")], "no bench line without a GPU"


def test_layout_flags():
    sys.path.insert(0, ROOT)
    import bench
    a = bench.parse_args([])
    assert (a.gpus, a.impl, a.config, a.shard_mode, a.replicas) == (1, "ours", "c2", "vector", 1)
    assert a.warmup >= 3
    a = bench.parse_args(["--shard-mode", "replica", "--gpus", "8"])
    assert a.shard_mode == "replica"
    with pytest.raises(SystemExit):
        bench.parse_args(["--shard-mode", "bogus"])
