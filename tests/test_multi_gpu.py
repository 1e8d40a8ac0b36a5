"""Multi-GPU test (needs >= 2 CUDA devices; skipped on the single-GPU box): the fused all-gather
(gpbt_log_posterior_scatter + PeerGather over torch symmetric memory) equals an NCCL all-gather."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_gather_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "tools", "peer_probe.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0 and "PEER_GATHER_PASS" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.gpu
def test_sharded_ensemble_sampler_two_ranks():
    """Replicated-walker sampler over 2 GPUs: same chain as the single-GPU sampler with the same seed,
    identical on both ranks (odd ensemble size, so the slices are padded)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29578",
           os.path.join(ROOT, "tools", "sharded_sampler_probe.py"), "1001", "10"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0 and "SHARDED_SAMPLER_PASS" in res.stdout and "RUN_MCMC_SHARDED_PASS" in res.stdout, \
        res.stdout[-2000:] + res.stderr[-2000:]
