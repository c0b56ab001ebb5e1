"""Parity of the row-sharded sweeps: sharded chains equal the single-GPU chains
and beta is bit-identical on every rank.

  * test_peer_window_exchange_two_ranks_one_gpu -- runs on ANY box with one GPU: two processes share
    cuda:0, the engine communicator is bl_comm_init_local (no NCCL), every exchange goes through the
    CUDA-IPC peer windows: the kernels of the production exchange (peer_publish in k_gram_reduce,
    peer_wait / peer_stage in k_beta_draw, k_peer_allreduce) on the production code path.
  * test_sharded_chain_equals_single_gpu_chain -- one rank per GPU over NCCL + NVLink peer windows,
    then ncclAllReduce on the same shards (needs >= 2 GPUs, skipped otherwise).
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, port, env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                           "--master-addr", "127.0.0.1", "--master-port", str(port),
                           os.path.join(ROOT, "tools", "check_multi_gpu.py")],
                          capture_output=True, text=True, timeout=1500, env=env)


def test_peer_window_exchange_two_ranks_one_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    out = _run(2, 29543, {"BL_MG_LOCAL": "1"})
    assert "MULTI_GPU_OK" in out.stdout, (out.stdout[-3000:], out.stderr[-3000:])
    assert "local mode (all ranks on cuda:0, no NCCL): True" in out.stdout


def test_sharded_chain_equals_single_gpu_chain():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run(2, 29541, {})
    assert "MULTI_GPU_OK" in out.stdout, (out.stdout[-3000:], out.stderr[-3000:])
