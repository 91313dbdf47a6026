"""Multi-GPU parity as a pytest (driver-visible): when the box shows at least two GPUs, tests/check_multigpu.py runs
under torch.distributed.run with one rank per GPU (2, and all of them when there are more) and must report that the
merged result of every sharding layout equals the oracle's search of the unsharded index.  Skipped on a one-GPU box;
the same exchange logic runs on the CPU with gloo in tests/test_distributed_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_search_equals_the_unsharded_oracle(world):
    n = _ngpus()
    if n < world:
        pytest.skip(f"{n} GPU(s) visible, {world} needed")
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29530 + world),
                        os.path.join(ROOT, "tests", "check_multigpu.py")],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=900)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    assert "MISMATCH" not in out
    assert out.count("merged result == oracle") >= 6, out[-3000:]
