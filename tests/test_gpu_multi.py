"""Multi-GPU parity (needs >= 2 GPUs on the box, skipped otherwise): row-sharded logit Gibbs with
the in-stream NCCL all-reduce reproduces the single-GPU chain."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_chain_equals_single_gpu_chain():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541",
                          os.path.join(ROOT, "tools", "check_multi_gpu.py")],
                         capture_output=True, text=True, timeout=900)
    assert "MULTI_GPU_OK" in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])
