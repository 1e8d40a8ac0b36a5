"""Device-resident PTLMC iteration loop (csrc/ptlmc.cuh, gpbt_ptlmc_*) against its NumPy restatement
(oracle/ptlmc_oracle.py: the reference's loop, src/mcmc.py:623-671 and :679-693, on the device's Philox draws)."""
import pickle

import numpy as np
import pytest

from oracle import gp_oracle as orc
from oracle import ptlmc_oracle as pto
from tests import goldens
from tests.helpers import product_states

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c1():
    import gpbt_b200  # noqa: F401
    from gpbt_b200.device import DeviceChain
    g = goldens.load("c1_rbf")
    states, sts = product_states(g)
    ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
    logp = lambda X: orc.log_posterior(sts, X, g["lo"], g["hi"], g["y_exp"], g["cov_exp"])   # the CPU oracle's
    yield g, ch, logp
    ch.release()


def setup_run(g, n_hot, n_cold, maxtemp, seed, spread=0.04):
    from gpbt_b200.ptlmc import temperature_ladder
    rng = np.random.default_rng(seed)
    lo, hi = g["lo"], g["hi"]
    n, p = n_hot + n_cold, len(lo)
    theta = 0.5 * (lo + hi) + spread * (hi - lo) * rng.standard_normal((n, p))
    temps = temperature_ladder(n_hot, n_cold, maxtemp).ravel()
    cov = np.cov(theta.T)
    cov = 0.9 * cov + 0.1 * np.diag(np.diag(cov))
    W, V = np.linalg.eigh(cov)
    return theta, temps, V @ np.diag(np.sqrt(W)) @ V.T


@pytest.mark.parametrize("n_hot,n_cold,n_tune,n_keep", [(8, 4, 25, 12), (50, 21, 12, 6), (0 + 1, 1, 11, 5)])
def test_device_loop_follows_the_oracle(c1, n_hot, n_cold, n_tune, n_keep):
    """trajectories, accept pattern, step-size tuning and exchange order of the device loop = the NumPy
    restatement fed with the same Philox draws, its log-posterior the CPU oracle's"""
    from gpbt_b200.ptlmc import DevicePTLMC
    g, ch, logpost = c1
    theta, temps, root = setup_run(g, n_hot, n_cold, 15.0, seed=n_hot)
    theta[1, 0] = g["hi"][0] - 1e-3            # a chain at the edge of the box: proposals leave it (lp = -inf)
    want = pto.run(logpost, theta, temps, root, n_hot, n_tune, n_keep, seed=77)
    dev = DevicePTLMC(ch, temps, root, n_hot, goal=0.25, seed=77)
    dev.set_state(theta, tau=-1.0)
    dev.run(n_tune, n_keep, n_steps=7)          # piecewise stepping continues the same run
    dev.run(n_tune, n_keep, n_steps=n_tune + n_keep - 7)
    got = dev.read()
    dev.close()
    assert want["takes"].any() and not want["takes"].all()
    assert got["accepted"] == want["accepted"]
    assert abs(got["tau"] - want["tau"]) <= 1e-12
    np.testing.assert_allclose(got["theta"], want["theta"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(got["state"], want["state"], rtol=0, atol=1e-9)
    fin = np.isfinite(want["lp"])
    assert np.array_equal(np.isfinite(got["lp"]), fin) and np.max(np.abs(got["lp"][fin] - want["lp"][fin])) <= 1e-7
    lo, hi = g["lo"], g["hi"]
    flat = got["theta"].reshape(-1, len(lo))
    assert np.all((flat > lo) & (flat < hi))    # nothing recorded outside the box


def test_exchange_with_thousands_of_chains(c1):
    """3000 chains: most 32-slot windows of a sweep are conflict free and are applied by all lanes at once, the
    others are replayed in order -- the result is the sequential sweep's"""
    from gpbt_b200.ptlmc import DevicePTLMC
    g, ch, logpost = c1
    theta, temps, root = setup_run(g, 2900, 100, 40.0, seed=3, spread=0.08)
    want = pto.run(logpost, theta, temps, root, 2900, 2, 1, seed=5)
    dev = DevicePTLMC(ch, temps, root, 2900, seed=5)
    dev.set_state(theta)
    dev.run(2, 1)
    got = dev.read()
    dev.close()
    np.testing.assert_allclose(got["state"], want["state"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(got["theta"], want["theta"], rtol=0, atol=1e-9)
    assert not np.allclose(got["state"], theta)


def test_argument_errors(c1):
    from gpbt_b200 import _lib
    from gpbt_b200.ptlmc import DevicePTLMC
    g, ch, _ = c1
    theta, temps, root = setup_run(g, 3, 2, 5.0, seed=1)
    with pytest.raises(ValueError):
        DevicePTLMC(ch, temps, root[:3, :3], 3)
    with pytest.raises(_lib.GpbtError):
        DevicePTLMC(ch, temps, root, 5)            # no T = 1 chain left
    dev = DevicePTLMC(ch, temps, root, 3, seed=1)
    with pytest.raises(_lib.GpbtError):
        dev.run(4, 4)                              # no state yet
    with pytest.raises(ValueError):
        dev.set_state(theta[:-1])
    dev.set_state(theta)
    dev.run(4, 4, n_steps=3)
    with pytest.raises(_lib.GpbtError):
        dev.run(5, 4, n_steps=1)                   # another run's shape while one is in progress
    with pytest.raises(_lib.GpbtError):
        dev.run(4, 4, n_steps=6)                   # more steps than the run has
    dev.close()
    with pytest.raises(RuntimeError):
        dev.run(4, 4)


def test_run_mcmc_ptlmc_device_loop(tmp_path):
    """Chain.run_MCMC_PTLMC(sampler="device"): host start-up stage (ranking, L-BFGS-B), device iteration loop;
    chain file layout [nwalkers, nsteps, ndim], T = 1 chains inside the box, in high-posterior territory, and
    reproducible from (NumPy seed, seed)"""
    from gpbt_b200 import synthetic
    from gpbt_b200.mcmc import Chain
    g = goldens.load("c1_rbf")
    states, _ = product_states(g)
    (tmp_path / "mcmc").mkdir()
    paths = synthetic.write_fixture(str(tmp_path), p=5, n=8, m=50)
    ch = Chain(mcmc_path=str(tmp_path / "mcmc" / "pt.pkl"), expdata_path=paths["exp"], model_parafile=paths["par"])
    ch.emuList = states
    chains = []
    for _ in range(2):
        np.random.seed(1)
        ch.run_MCMC_PTLMC(nsteps=30, nwalkers=4, ntemps=6, maxtemp=20, nstartparameters=80, sampler="device", seed=9)
        with open(ch.mcmc_path, "rb") as fh:
            chains.append(pickle.load(fh)["chain"])
    chain = chains[0]
    assert chain.shape == (4, 30, 5)
    np.testing.assert_array_equal(chains[0], chains[1])
    flat = chain.reshape(-1, 5)
    assert np.all((flat > ch.min) & (flat < ch.max))
    lp = ch.log_posterior(flat)
    assert np.all(np.isfinite(lp)) and np.median(lp) > np.median(ch.log_posterior(ch.random_pos(200)))
    assert len(np.unique(flat[:, 0])) > 8          # the chains move
    with pytest.raises(ValueError):
        ch.samplerPTLMC(ch.log_posterior, ch.random_pos, sampler="gpu")
