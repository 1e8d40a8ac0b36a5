"""Parallel-tempering driver (gpbt_b200.ptlmc, Chain.samplerPTLMC / tempexchange) against golden
chains produced by the UNMODIFIED reference on a toy log-posterior with the same NumPy seed
(tests/golden/make_golden_ptlmc.py -> ptlmc_toy.npz).  CPU only: the sampler is host code that calls
whatever log-posterior it is given."""
import importlib.util
import os

import numpy as np
import pytest

import gpbt_b200  # noqa: F401
from gpbt_b200 import ptlmc

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden_ptlmc", os.path.join(HERE, "golden", "make_golden_ptlmc.py"))


@pytest.fixture(scope="module")
def toy():
    # the toy target lives in the generator script (its import of the reference happens in main() only)
    mod = importlib.util.module_from_spec(spec)
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    spec.loader.exec_module(mod)
    return mod, np.load(os.path.join(HERE, "golden", "ptlmc_toy.npz"))


def test_temp_exchange_matches_reference(toy):
    mod, g = toy
    np.random.seed(100)
    order = ptlmc.temp_exchange(g["ex_lp"], g["ex_temps"], iters=4)
    np.testing.assert_array_equal(order, g["ex_order"])
    assert sorted(order) == list(range(9))
    np.random.seed(100)
    np.testing.assert_array_equal(ptlmc.temp_exchange_python(g["ex_lp"], g["ex_temps"], iters=4), g["ex_order"])
    # a ladder long enough for many dependent swaps: helper and Python loop agree draw for draw
    rng = np.random.default_rng(5)
    lp, temps = rng.normal(size=(700, 1)) * 40, ptlmc.temperature_ladder(500, 200, 50.0)
    np.random.seed(8)
    a = ptlmc.temp_exchange(lp, temps, iters=5)
    state = np.random.get_state()[1][:4].copy()
    np.random.seed(8)
    b = ptlmc.temp_exchange_python(lp, temps, iters=5)
    np.testing.assert_array_equal(a, b)
    assert np.array_equal(state, np.random.get_state()[1][:4]) and (a != np.arange(700)).sum() > 100


def test_chain_matches_reference_without_gradient(toy):
    mod, g = toy
    np.random.seed(20261018)
    out = ptlmc.sampler_ptlmc(mod.toy_logpost, mod.draw, theta0=None, numtemps=5, numchain=3, sampperchain=25,
                              maxtemp=12, nstartparameters=80)["theta"]
    assert out.shape == (3, 25, 3)
    np.testing.assert_allclose(out, g["theta"], rtol=1e-10, atol=1e-12)
    assert np.all((out > mod.LO) & (out < mod.HI))


def test_chain_matches_reference_with_gradient(toy):
    mod, g = toy
    np.random.seed(7)
    out = ptlmc.sampler_ptlmc(mod.toy_with_grad, mod.draw, theta0=None, numtemps=4, numchain=2, sampperchain=15,
                              maxtemp=8, nstartparameters=60)["theta"]
    np.testing.assert_allclose(out, g["theta_grad"], rtol=1e-10, atol=1e-12)


def test_ladder_and_target_conventions():
    t = ptlmc.temperature_ladder(4, 2, 16.0)
    assert t.shape == (6, 1) and t[0, 0] == pytest.approx(16.0) and np.all(t[4:] == 1.0) and np.all(np.diff(t[:4, 0]) < 0)
    with pytest.raises(ValueError):
        ptlmc._Target(lambda X: (np.zeros(len(X)), np.zeros((len(X), 2)), 0), np.zeros((2, 3)), np.zeros(3))
    with pytest.raises(ValueError):
        ptlmc._Target(lambda X: (np.zeros((len(X), 1)), np.zeros((len(X), 2))), np.zeros((2, 3)), np.zeros(3))


def test_samples_a_gaussian(toy):
    """longer run: the T = 1 chains reproduce the toy posterior's mean and spread"""
    mod, _ = toy
    np.random.seed(3)
    out = ptlmc.sampler_ptlmc(mod.toy_logpost, mod.draw, numtemps=8, numchain=8, sampperchain=600, maxtemp=20,
                              nstartparameters=200)["theta"].reshape(-1, 3)
    sd = np.sqrt(np.diag(mod.COV))
    assert np.all(np.abs(out.mean(0) - mod.MEAN) < 0.25 * sd)
    assert np.all(np.abs(out.std(0) / sd - 1) < 0.25)
