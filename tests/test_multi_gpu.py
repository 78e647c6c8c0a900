"""N > 1 on real GPUs over NCCL ("multi-GPU tests that compare N-GPU output with 1-GPU output bit for
bit"): launches tests/mgpu_worker.py under torchrun on 2 GPUs and on all visible GPUs.  Skipped on a one-GPU box (NCCL
refuses two ranks on one device); the gloo tests of tests/test_distributed_cpu.py cover the host logic there, and
bench.py --gpus N repeats the same equality checks inside the driver's scaling run (its JSON line: "verified")."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worlds():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return sorted({w for w in (2, n) if 2 <= w <= n})


@pytest.mark.parametrize("world", [2, 4, 8])
def test_n_gpu_results_equal_one_gpu_results(world):
    if world not in _worlds() and not (world == 4 and torch.cuda.is_available() and torch.cuda.device_count() >= 4):
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(REPO, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, cwd=REPO, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and ("MGPU-OK %d" % world) in r.stdout, r.stdout[-4000:]
