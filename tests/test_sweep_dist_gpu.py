"""-m gpu, world size 2: the product's own collective (mtg_nccl_init + mtg_argmin_allgather /
mtg_best_allgather over NCCL) against torch.distributed's gather and against a serial scan of all
costs, one process per GPU under torchrun. Skipped on a single-GPU box (the driver's 1-GPU test
tier); run with `gpurun --gpus 2`. The gloo / world-size-2 CPU test of the same host logic is
tests/test_sweep_dist_cpu.py."""
import json
import os
import socket
import subprocess
import sys

import pytest

from gpu_util import require_cuda

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("total", [200_000, 100_001])
def test_library_collective_world_2(total):
    torch = require_cuda()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "nccl_argmin_check.py"), "--total", str(total)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2 and out["all_ranks_agree"]
