"""Multi-GPU parity through the CUDA path (needs >= 2 GPUs; skipped otherwise): tests/dist_gpu_check.py under torchrun --
data parallel == single process for both schedules (peer-memory kernel, NCCL all-reduce), replicas bitwise equal, sharded
top-k == unsharded."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_and_sharded_scoring_on_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "dist_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST CHECK PASSED" in r.stdout, r.stdout[-4000:] + r.stderr[-4000:]
