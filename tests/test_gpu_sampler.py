"""GPU tests of the device-resident ensemble sampler (csrc/ensemble.cuh, gpbt_ensemble_*): draw-for-draw
parity with the NumPy restatement of the stretch move (oracle/ensemble_oracle.py) whose log-posterior
is the oracle's, the device Philox streams, graph replay vs eager enqueue, and the reference's
run_mcmc recipe on top of it."""
import pickle

import numpy as np
import pytest

from oracle import ensemble_oracle as eo
from oracle import gp_oracle as orc
from tests import goldens
from tests.helpers import ABS_LP, product_states

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c1():
    from gpbt_b200.device import DeviceChain
    g = goldens.load("c1_rbf")
    states, sts = product_states(g)
    dc = DeviceChain(states, g["lo"], g["hi"], g["y_exp"], g["cov_exp"])
    logp = lambda X: orc.log_posterior(sts, X, g["lo"], g["hi"], g["y_exp"], g["cov_exp"])
    yield g, dc, logp
    dc.release()


def start(g, nw, seed=5):
    rng = np.random.default_rng(seed)
    mid, half = 0.5 * (g["lo"] + g["hi"]), 0.5 * (g["hi"] - g["lo"])
    return mid + 0.6 * half * rng.uniform(-1, 1, (nw, len(mid)))


@pytest.mark.parametrize("nw,use_graph", [(16, True), (16, False), (15, True), (64, True)])
def test_trajectory_matches_oracle_draw_for_draw(c1, nw, use_graph):
    """Same random draws -> same proposals (bit for bit: the kernel rounds c - (c - s) z operation by
    operation), same accept decisions, log-posteriors within the 1e-8 budget, for even and odd
    ensembles; wide proposals (a = 2.5) push some walkers out of the box, which must be rejected."""
    from gpbt_b200.sampler import DeviceEnsembleSampler
    g, dc, logp = c1
    steps, n0 = 25, (nw + 1) // 2
    rng = np.random.default_rng(nw)
    u = rng.random((steps, 2, n0, 2))
    partner = np.stack([rng.integers(0, nw - n0, (steps, n0)), rng.integers(0, n0, (steps, n0))], axis=1).astype(np.int32)
    perm = np.stack([rng.permutation(nw) for _ in range(steps)]).astype(np.int32)
    x0 = start(g, nw)
    s = DeviceEnsembleSampler(nw, x0.shape[1], dc, a=2.5, seed=1, use_graph=use_graph)
    s.set_state(x0)
    lp0 = s.get_state().log_prob
    assert np.max(np.abs(lp0 - logp(x0))) <= ABS_LP
    s.advance(steps, u=u, partner=partner, perm=perm)
    chain, lps, acc = eo.stretch_run(logp, x0, logp(x0), u, partner, perm, a=2.5)
    got = s.get_chain()
    assert got.shape == (steps, nw, x0.shape[1])
    np.testing.assert_array_equal(got, chain)
    got_lp = s.get_log_prob()
    assert np.max(np.abs(got_lp - lps)) <= ABS_LP
    np.testing.assert_array_equal(s.n_accepted, acc)
    assert 0 < acc.sum() < steps * nw                      # both outcomes occurred
    inside = np.all((got > g["lo"]) & (got < g["hi"]), axis=2)
    assert inside.all() and np.all(np.isfinite(got_lp))    # out-of-box proposals never get accepted
    st = s.get_state()
    np.testing.assert_array_equal(st.coords, chain[-1])
    s.close()


@pytest.mark.parametrize("nw", [12, 2101])
def test_device_philox_streams(c1, nw):
    """Production mode (no host draws): the device's counter-based streams equal the Python Philox
    restatement, across two consecutive runs (the step counter continues) and after reset() -- for
    a small ensemble (one fused CTA per piece of a step) and one above that path's limit (separate
    key / rank / propose / accept / record kernels, odd size)."""
    from gpbt_b200.sampler import DeviceEnsembleSampler
    g, dc, logp = c1
    seed = 0x1234567887654321
    x0 = start(g, nw, 9)
    s = DeviceEnsembleSampler(nw, x0.shape[1], dc, seed=seed)
    s.set_state(x0)
    s.advance(7)
    s.advance(6)
    u, partner, perm = eo.philox_streams(seed, 0, 13, nw)
    for k in range(13):
        assert sorted(perm[k]) == list(range(nw))
    chain, lps, acc = eo.stretch_run(logp, x0, logp(x0), u, partner, perm)
    np.testing.assert_array_equal(s.get_chain(), chain)
    assert s.iteration == 13 and np.max(np.abs(s.get_log_prob() - lps)) <= ABS_LP
    # reset: history and counters start again at step 0, the walkers stay
    s.reset()
    assert s.iteration == 0 and s.get_chain().shape[0] == 0
    s.advance(3)
    u, partner, perm = eo.philox_streams(seed, 0, 3, nw)
    chain2, _, acc2 = eo.stretch_run(logp, chain[-1], lps[-1], u, partner, perm)
    np.testing.assert_array_equal(s.get_chain(), chain2)
    np.testing.assert_array_equal(s.n_accepted, acc2)
    s.close()


def test_graph_replay_equals_eager_and_fixed_split(c1):
    from gpbt_b200.sampler import DeviceEnsembleSampler
    g, dc, logp = c1
    nw = 32
    x0 = start(g, nw, 2)
    out = []
    for use_graph in (True, False):
        s = DeviceEnsembleSampler(nw, x0.shape[1], dc, seed=77, use_graph=use_graph, randomize_split=False)
        s.set_state(x0)
        # an unrelated, larger call in between reallocates the chain's workspaces: the captured graph
        # has to notice and be rebuilt
        s.advance(5)
        dc.log_target(np.repeat(x0, 40, axis=0), -np.inf)
        s.advance(5)
        out.append((s.get_chain(), s.get_log_prob()))
        s.close()
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])
    u, partner, _ = eo.philox_streams(77, 0, 10, nw)
    perm = np.tile(eo.fixed_split(nw), (10, 1))
    chain, _, _ = eo.stretch_run(logp, x0, logp(x0), u, partner, perm)
    np.testing.assert_array_equal(out[0][0], chain)


def test_posterior_moments_agree_with_host_sampler(c1):
    """Statistical check: a long device run and a long host run of the same move (tests/fake_emcee.py
    through the GPU log-posterior) sample the same posterior -- means within 0.15 posterior standard
    deviations, standard deviations within 15 %."""
    from gpbt_b200.sampler import DeviceEnsembleSampler
    from tests import fake_emcee
    g, dc, _ = c1
    nw, steps, burn = 48, 3000, 500
    x0 = start(g, nw, 11)
    s = DeviceEnsembleSampler(nw, x0.shape[1], dc, seed=2026)
    s.run_mcmc(x0, steps)
    dev = s.get_chain(discard=burn, flat=True)
    af = s.acceptance_fraction
    s.close()
    host = fake_emcee.EnsembleSampler(nw, x0.shape[1], lambda X: dc.log_target(X, -np.inf),
                                      pool=type("P", (), {"map": staticmethod(lambda f, a: f(a))}), seed=3)
    for _ in host.sample(x0, iterations=steps):
        pass
    ref = host.get_chain()[burn:].reshape(-1, x0.shape[1])
    sd = ref.std(axis=0)
    assert np.all(np.abs(dev.mean(axis=0) - ref.mean(axis=0)) <= 0.15 * sd), (dev.mean(0), ref.mean(0), sd)
    assert np.all(np.abs(dev.std(axis=0) / sd - 1.0) <= 0.15), (dev.std(0), sd)
    assert 0.1 < af.mean() < 0.9 and abs(af.mean() - host.acceptance_fraction.mean()) < 0.05


def test_run_mcmc_on_device(tmp_path):
    """Chain.run_mcmc with the device sampler: the reference's burn-in recipe (src/mcmc.py:366-405),
    chain file layout [walker, thinned step, dim], restart from the stored chain."""
    from gpbt_b200 import synthetic
    from gpbt_b200.mcmc import Chain
    g = goldens.load("c1_rbf")
    states, sts = product_states(g)
    (tmp_path / "mcmc").mkdir()
    paths = synthetic.write_fixture(str(tmp_path), p=5, n=8, m=50)
    ch = Chain(mcmc_path=str(tmp_path / "mcmc" / "chain.pkl"), expdata_path=paths["exp"], model_parafile=paths["par"])
    ch.emuList = states
    np.random.seed(0)
    ch.run_mcmc(nsteps=30, nburnsteps=20, nwalkers=16, nthin=3, seed=4, sampler="device")
    with open(ch.mcmc_path, "rb") as fh:
        chain = pickle.load(fh)["chain"]
    assert chain.shape == (16, 10, 5)
    flat = chain.reshape(-1, 5)
    assert np.all((flat > ch.min) & (flat < ch.max))
    lp = ch.log_posterior(flat)
    want = orc.log_posterior(sts, flat[:12], ch.min, ch.max, ch.expdata, ch.expdata_cov)
    assert np.all(np.isfinite(lp)) and np.max(np.abs(lp[:12] - want)) <= ABS_LP
    assert ch.acceptance_fraction_.shape == (16,) and ch.acceptance_fraction_.max() > 0
    ch.run_mcmc(nsteps=6, nburnsteps=20, nwalkers=16, nthin=3, seed=5, sampler="device")
    with open(ch.mcmc_path, "rb") as fh:
        again = pickle.load(fh)["chain"]
    assert again.shape == (16, 12, 5)
    np.testing.assert_array_equal(again[:, :10], chain)
    # same seeds -> the same chain, bit for bit
    ch2 = Chain(mcmc_path=str(tmp_path / "mcmc" / "chain2.pkl"), expdata_path=paths["exp"], model_parafile=paths["par"])
    ch2.emuList = states
    np.random.seed(0)
    ch2.run_mcmc(nsteps=30, nburnsteps=20, nwalkers=16, nthin=3, seed=4, sampler="device")
    np.testing.assert_array_equal(ch2.chain, chain)


def test_argument_errors(c1):
    from gpbt_b200 import _lib
    from gpbt_b200.device import DeviceChain
    from gpbt_b200.sampler import DeviceEnsembleSampler
    g = c1[0]
    dc = DeviceChain(product_states(g)[0], g["lo"], g["hi"], g["y_exp"], g["cov_exp"])
    with pytest.raises(ValueError):
        DeviceEnsembleSampler(8, 4, dc)                     # wrong ndim
    with pytest.raises(_lib.GpbtError):
        DeviceEnsembleSampler(8, 5, dc, a=1.0)              # a must exceed 1
    s = DeviceEnsembleSampler(8, 5, dc)
    with pytest.raises(RuntimeError):
        s.advance(1)                                        # no state yet
    with pytest.raises(ValueError):
        s.run_mcmc(np.tile(start(g, 1), (8, 1)), 2)         # identical walkers: degenerate start
    s.run_mcmc(np.tile(start(g, 1), (8, 1)), 2, skip_initial_state_check=True)
    assert np.ptp(s.get_chain(), axis=1).max() == 0.0       # c - (c - s) z with c == s: nobody can move
    dc.release()                                            # releasing the chain closes its samplers
    with pytest.raises(RuntimeError):
        s.advance(1)


def test_run_mcmc_ptlmc_on_gpu_posterior(tmp_path):
    """Chain.run_MCMC_PTLMC (src/mcmc.py:696-727): the parallel-tempering driver on the GPU
    log-posterior -- N = 1 and N = 2 probe calls, N = 1 calls from L-BFGS-B, one call of all chains
    per iteration; chain file layout [nwalkers, nsteps, ndim], T = 1 chains inside the box."""
    from gpbt_b200 import synthetic
    from gpbt_b200.mcmc import Chain
    g = goldens.load("c1_rbf")
    states, sts = product_states(g)
    (tmp_path / "mcmc").mkdir()
    paths = synthetic.write_fixture(str(tmp_path), p=5, n=8, m=50)
    ch = Chain(mcmc_path=str(tmp_path / "mcmc" / "pt.pkl"), expdata_path=paths["exp"], model_parafile=paths["par"])
    ch.emuList = states
    np.random.seed(1)
    ch.run_MCMC_PTLMC(nsteps=20, nwalkers=4, ntemps=6, maxtemp=20, nstartparameters=80)
    with open(ch.mcmc_path, "rb") as fh:
        chain = pickle.load(fh)["chain"]
    assert chain.shape == (4, 20, 5)
    flat = chain.reshape(-1, 5)
    assert np.all((flat > ch.min) & (flat < ch.max))
    lp = ch.log_posterior(flat)
    want = orc.log_posterior(sts, flat[:10], ch.min, ch.max, ch.expdata, ch.expdata_cov)
    assert np.all(np.isfinite(lp)) and np.max(np.abs(lp[:10] - want)) <= ABS_LP
    # the optimiser and the tempered chains have moved to high-posterior territory
    assert np.median(lp) > np.median(ch.log_posterior(ch.random_pos(200)))


@pytest.mark.parametrize("name", ["c1_matern", "c1_logexp", "c1_nopca", "c1_multi", "odd_shape", "p20_trafo"])
def test_sampler_on_every_chain_kind(name):
    """run_mcmc's default sampler has to work on whatever path the chain takes -- diagonal
    (exp/diag emulators), dense Cholesky (no-PCA), several emulators, the parameter-function
    pre-transform -- under graph replay: trajectories against the oracle from the device's Philox draws."""
    from gpbt_b200.device import DeviceChain
    from gpbt_b200.sampler import DeviceEnsembleSampler
    if name not in goldens.available():
        pytest.skip("golden %s not present" % name)
    g = goldens.load(name)
    states, sts = product_states(g)
    dc = DeviceChain(states, g["lo"], g["hi"], g["y_exp"], g["cov_exp"])
    logp = lambda X: orc.log_posterior(sts, X, g["lo"], g["hi"], g["y_exp"], g["cov_exp"])
    nw, steps, seed = 2 * len(g["lo"]) + 3, 8, 42
    x0 = start(g, nw, 3)
    s = DeviceEnsembleSampler(nw, x0.shape[1], dc, seed=seed)
    s.set_state(x0)
    s.advance(steps)
    u, partner, perm = eo.philox_streams(seed, 0, steps, nw)
    chain, lps, acc = eo.stretch_run(logp, x0, logp(x0), u, partner, perm)
    np.testing.assert_array_equal(s.get_chain(), chain)
    got = s.get_log_prob()
    fin = np.isfinite(lps)
    assert np.array_equal(np.isfinite(got), fin) and np.max(np.abs(got[fin] - lps[fin])) <= 2 * ABS_LP
    np.testing.assert_array_equal(s.n_accepted, acc)
    s.close()
    dc.release()


@pytest.mark.parametrize("nw", [2, 3, 5, 2048, 2049])
def test_ensemble_size_boundaries(c1, nw):
    """Smallest ensembles (one walker per set), and both sides of the switch from key ranking to the
    keyed-bijection split (2048 / 2049 walkers) -- device Philox draws against the oracle's."""
    from gpbt_b200.sampler import DeviceEnsembleSampler
    g, dc, logp = c1
    steps, seed = 3, 2026
    x0 = start(g, nw, 21)
    s = DeviceEnsembleSampler(nw, x0.shape[1], dc, seed=seed)
    s.set_state(x0)
    s.advance(steps)
    u, partner, perm = eo.philox_streams(seed, 0, steps, nw)
    for k in range(steps):
        assert sorted(perm[k]) == list(range(nw))
    chain, lps, acc = eo.stretch_run(logp, x0, logp(x0), u, partner, perm)
    np.testing.assert_array_equal(s.get_chain(), chain)
    np.testing.assert_array_equal(s.n_accepted, acc)
    s.close()


def test_graph_replay_over_the_stepped_cholesky():
    """A chain without low-rank factors (dense path) with m > 80 and half-ensembles of 256+ walkers runs
    the panel-synchronous Cholesky (19+ dependent launches with programmatic dependent launch) inside the
    captured step: graph replay and eager launches must give the same chain, and both the oracle's."""
    from gpbt_b200 import synthetic
    from gpbt_b200.device import DeviceChain
    from gpbt_b200.sampler import DeviceEnsembleSampler
    from gpbt_b200.state import EmulatorState
    arr = synthetic.untrained_state_arrays(4, 40, 96, 5)
    st = EmulatorState.from_arrays(**arr, keep_L=True)
    od = st.oracle_dict()
    lo, hi = synthetic.box(4)
    y = orc.emulator_predict(od, (0.5 * (lo + hi))[None, :], False)[0]
    cov_exp = np.diag((0.05 * np.abs(y)) ** 2 + 1e-3)
    dc = DeviceChain([st], lo, hi, y, cov_exp, lowrank=False)
    logp = lambda X: orc.log_posterior([od], X, lo, hi, y, cov_exp)
    nw, steps, seed = 600, 3, 11
    rng = np.random.default_rng(1)
    x0 = 0.5 * (lo + hi) + 0.3 * (hi - lo) * rng.uniform(-1, 1, (nw, 4))
    out = []
    for use_graph in (True, False):
        s = DeviceEnsembleSampler(nw, 4, dc, seed=seed, use_graph=use_graph)
        s.set_state(x0)
        s.advance(steps)
        out.append((s.get_chain(), s.get_log_prob()))
        s.close()
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])
    u, partner, perm = eo.philox_streams(seed, 0, steps, nw)
    chain, lps, _ = eo.stretch_run(logp, x0, logp(x0), u, partner, perm)
    np.testing.assert_array_equal(out[0][0], chain)
    assert np.max(np.abs(out[0][1] - lps)) <= ABS_LP
    dc.release()
