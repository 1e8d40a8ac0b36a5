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


@pytest.mark.gpu
def test_single_process_fanout_matches_one_gpu():
    """Chain.log_posterior from ONE process over all visible GPUs (gpbt_fanout_*: replica + worker thread
    + pinned staging per GPU) returns what a single GPU returns (to 1e-10: a smaller per-GPU batch may
    take a narrower walker tile, i.e. another summation order) -- ragged batch sizes, forced and
    automatic device counts, out-of-bounds rows, and the dense path."""
    import numpy as np
    import gpbt_b200  # noqa: F401
    from gpbt_b200 import _lib
    from gpbt_b200.device import DeviceChain
    from tests import goldens
    from tests.helpers import product_states
    n = _lib.lib.gpbt_device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    g = goldens.load("c1_rbf")
    states, _ = product_states(g)
    devs = list(range(n))
    ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"], devices=devs)
    rng = np.random.default_rng(5)
    for N in (1, 37, 2048, 5000, 70001):
        X = rng.uniform(g["lo"], g["hi"], (N, len(g["lo"])))
        X[::53, 0] = g["hi"][0] + 1.0
        one = ch.log_target(X, -np.inf, max_devices=1)
        assert ch.last_devices_used == 1
        fin = np.isfinite(one)
        assert not fin[::53].any() and fin[1::53].all()
        for md in (None, 2, n):
            got = ch.log_target(X, -np.inf, max_devices=md)
            assert np.array_equal(np.isfinite(got), fin) and np.max(np.abs(got[fin] - one[fin]), initial=0.0) <= 1e-10
            if md and N >= 16 * md:
                assert ch.last_devices_used == md
        if N <= 5000:
            a = ch.log_target(X, -1e300, path="dense", max_devices=n)
            b = ch.log_target(X, -1e300, path="dense", max_devices=1)
            assert np.max(np.abs(a - b)) <= 1e-9 and np.all(a[::53] == -1e300)
    auto = ch.log_target(rng.uniform(g["lo"], g["hi"], (4096, len(g["lo"]))), -np.inf)
    assert ch.last_devices_used == min(n, 16) and auto.shape == (4096,)
    ch.release()
