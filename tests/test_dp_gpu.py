"""GPU, 2 ranks over NCCL (skipped on a one-GPU box): the data-parallel fused step - bucketed gradient all-reduce
launched from inside the backward passes on a communication stream, eager and captured in the CUDA graph - gives the
hand-summed gradient of N single-GPU replicas on the same shards and keeps the replicas' parameters bit-identical."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_dp_step_matches_summed_replicas():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tests", "multigpu", "dp_step_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DP_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
