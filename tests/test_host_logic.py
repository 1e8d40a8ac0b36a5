"""CPU tests of the host-side logic that feeds the kernels: the exact low-rank factors
(device.lowrank_factors) against the dense mvn_loglike of the oracle, state extraction from trained
sklearn objects, the parameter-function curves, and pickling."""
import os

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import gp_oracle as orc
from tests import goldens
from tests.helpers import product_states


def _fake_states(rng, m_list, q_list):
    import gpbt_b200  # noqa: F401
    from types import SimpleNamespace
    out = []
    for m, q in zip(m_list, q_list):
        B = rng.normal(size=(m + 3, m))
        out.append(SimpleNamespace(m=m, q=q, A=rng.normal(size=(q, m)), mu=rng.normal(size=m) + 3.0,
                                   Ctrunc=B.T @ B / m + 0.05 * np.eye(m), no_pca=False, exp_diag=False))
    return out


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10 ** 6), n_emu=st.integers(1, 3), full=st.booleans(), scale=st.floats(0.01, 30.0))
def test_lowrank_identity_matches_dense(seed, n_emu, full, scale):
    """log det / quadratic form through (R, c0, s_perp, logdetF) == dpotrf/dpotrs on the assembled
    covariance (src/mcmc.py:23-65), for random chains, with a diagonal or a full experimental
    covariance, and PC variances spanning small to large (scale)."""
    from gpbt_b200.device import lowrank_factors
    rng = np.random.default_rng(seed)
    m_list = [int(rng.integers(3, 9)) for _ in range(n_emu)]
    q_list = [int(rng.integers(1, mm)) for mm in m_list]
    states = _fake_states(rng, m_list, q_list)
    M, Q = sum(m_list), sum(q_list)
    y_exp = rng.normal(size=M) + 3.0
    cov_exp = np.diag(rng.uniform(0.01, 0.2, M))
    if full:
        G = 0.1 * rng.normal(size=(M, 2))
        cov_exp = cov_exp + G @ G.T
    lr = lowrank_factors(states, y_exp, cov_exp)
    assert lr["R"].shape == (Q, Q) and np.allclose(lr["R"], np.triu(lr["R"]))
    for _ in range(3):
        z = rng.normal(size=Q)
        v = scale * rng.uniform(1e-4, 1.0, Q)
        # dense reference: block-diagonal model covariance + experimental covariance
        C = cov_exp.copy()
        mean = np.empty(M)
        qo = mo = 0
        for s in states:
            C[mo:mo + s.m, mo:mo + s.m] += (s.A.T * v[qo:qo + s.q]) @ s.A + s.Ctrunc
            mean[mo:mo + s.m] = z[qo:qo + s.q] @ s.A + s.mu
            qo += s.q
            mo += s.m
        want = orc.mvn_loglike(mean - y_exp, C)
        c = lr["R"] @ z + lr["c0"]
        S = np.eye(Q) + (lr["R"] * v) @ lr["R"].T
        Ls = np.linalg.cholesky(S)
        t = np.linalg.solve(Ls, c)
        got = -0.5 * (lr["s_perp"] + t @ t) - np.log(np.diag(Ls)).sum() - lr["logdetF_half"]
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want))


def test_state_from_trained_matches_golden_arrays(tmp_path):
    """EmulatorState.from_trained on this package's Emulator reproduces what it was trained to:
    L Linv = I, oracle predictions from the extracted state equal sklearn's own per-GP predict."""
    import gpbt_b200  # noqa: F401
    from gpbt_b200 import synthetic
    from gpbt_b200.emulator import Emulator
    paths = synthetic.write_fixture(str(tmp_path), p=3, n=40, m=10)
    for kind in ("RBF", "Matern"):
        emu = Emulator(training_set_path=paths["train"], parameter_file=paths["par"], npc=4)
        emu.trainEmulator([True] * emu.nev, kernel_type=kind)
        stt = emu.state
        assert stt.kind == kind and stt.p == 3 and stt.n == 40 and stt.q == 4 and stt.m == 10
        for j in range(stt.q):
            assert np.max(np.abs(stt.Linv[j] @ stt.L[j] - np.eye(stt.n))) < 1e-10
        X = synthetic.walkers(3, 20, seed=4, frac_outside=0.0)
        zm, zv = orc.pc_predict(stt.oracle_dict(), X)
        for j, gp in enumerate(emu.gps):
            mu_j, cov_j = gp.predict(X, return_cov=True)
            assert np.max(np.abs(zm[:, j] - mu_j)) < 1e-9 and np.max(np.abs(zv[:, j] - cov_j.diagonal())) < 1e-9
        # the pickled emulator carries no device handle and still knows its state
        import dill
        emu2 = dill.loads(dill.dumps(emu))
        assert emu2._device is None and emu2.state.q == 4


def test_curves_match_reference_piecewise_definitions():
    """curve_on_grid against a literal scalar transcription of src/emulator.py:100-124."""
    import gpbt_b200  # noqa: F401
    from gpbt_b200.emulator import curve_on_grid

    def zeta(a, T):
        sig = a[3] if T < a[1] else a[2]
        return a[0] * np.exp(-(T - a[1]) ** 2. / (2. * sig ** 2.))

    def eta(a, x):
        if 0. < x <= 0.2:
            return a[0] + (a[1] - a[0]) * (x / 0.2)
        if 0.2 < x < 0.4:
            return a[1] + (a[2] - a[1]) * ((x - 0.2) / 0.2)
        return a[2]

    def yl(a, y):
        if 0. < y <= 2.:
            return a[0] * (y / 2.)
        if 2. < y < 4.:
            return a[0] + (a[1] - a[0]) * ((y - 2.) / 2.)
        return a[1] + (a[2] - a[1]) * ((y - 4.) / 2.)

    rng = np.random.default_rng(3)
    for kind, fn, nargs, grid in ((0, zeta, 4, np.linspace(0, .5, 100)), (1, eta, 3, np.linspace(0, .6, 100)),
                                  (2, yl, 3, np.linspace(0, 6.2, 100))):
        P = rng.uniform(0.05, 1.0, (7, nargs))
        got = curve_on_grid(kind, P, grid)
        want = np.array([[fn(row, x) for x in grid] for row in P])
        assert np.allclose(got, want, rtol=0, atol=1e-15)
        # the oracle's vectorised copies agree too
        o = orc._CURVES[kind](*[P[:, i][:, None] for i in range(nargs)], grid[None, :])
        assert np.allclose(o, want, rtol=0, atol=1e-15)


def test_golden_states_are_consistent():
    """alpha_ and L_ in the golden files belong together: L L^T alpha reproduces sklearn's training
    targets up to the PCA whitening (unit variance) -- guards against mixing up fixture arrays."""
    g = goldens.load("c1_rbf")
    states, sts = product_states(g)
    s = sts[0]
    y = np.stack([s["L"][j] @ (s["L"][j].T @ s["alpha"][j]) for j in range(len(s["c"]))])
    assert np.all(np.abs(y.std(axis=1) - 1.0) < 0.05) and np.all(np.abs(y.mean(axis=1)) < 1e-8)
    assert states[0].device_bytes() > 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference tree exists only in the build container")
def test_state_extraction_from_a_reference_pickle(tmp_path):
    """`Chain.loadEmulator` accepts dill pickles made by the REFERENCE's Emulator class: train one
    with the unmodified reference, dill round trip, EmulatorState.from_trained, and the oracle on the
    extracted state reproduces the reference's own predict / log_posterior."""
    import subprocess
    import sys
    import textwrap
    import gpbt_b200  # noqa: F401
    from gpbt_b200 import synthetic
    paths = synthetic.write_fixture(str(tmp_path), p=3, n=30, m=8)
    # the reference is trained in a subprocess (its package is called `src`, and importing
    # src.mcmc needs stub emcee / pocomc modules)
    script = textwrap.dedent("""
        import os, sys, types, dill, numpy as np
        os.environ["WORKDIR"] = %r; os.environ["LOGLEVEL"] = "error"
        sys.path.insert(0, "/root/reference")
        em = types.ModuleType("emcee"); em.EnsembleSampler = type("EnsembleSampler", (), {"__init__": lambda s, *a, **k: None})
        pm = types.ModuleType("pocomc"); pm.Prior = object; pm.Sampler = object
        sys.modules["emcee"] = em; sys.modules["pocomc"] = pm
        from src.emulator import Emulator
        from src.mcmc import Chain
        emu = Emulator(training_set_path=%r, parameter_file=%r, npc=3)
        emu.trainEmulator([True] * emu.nev, kernel_type="Matern")
        dill.dump(emu, open(%r, "wb"))
        os.makedirs(os.path.join(%r, "mcmc"), exist_ok=True)
        ch = Chain(mcmc_path=os.path.join(%r, "mcmc", "c.pkl"), expdata_path=%r, model_parafile=%r)
        ch.loadEmulator([%r])
        X = np.load(%r)
        mean, cov = emu.predict(X, return_cov=True, extra_std=np.zeros(len(X)))
        np.savez(%r, mean=mean, cov=cov, lp=ch.log_posterior(X), y=ch.expdata, c=ch.expdata_cov, lo=ch.min, hi=ch.max)
    """) % (str(tmp_path), paths["train"], paths["par"], str(tmp_path / "ref_emu.pkl"), str(tmp_path), str(tmp_path),
            paths["exp"], paths["par"], str(tmp_path / "ref_emu.pkl"), str(tmp_path / "X.npy"), str(tmp_path / "ref_out.npz"))
    X = synthetic.walkers(3, 25, seed=2)
    np.save(tmp_path / "X.npy", X)
    res = subprocess.run([sys.executable, "-W", "ignore", "-c", script], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    ref = np.load(tmp_path / "ref_out.npz")
    # unpickling needs the reference's package on the path, exactly as in a user's environment
    import dill
    sys.path.insert(0, "/root/reference")
    try:
        with open(tmp_path / "ref_emu.pkl", "rb") as fh:
            emu = dill.load(fh)
    finally:
        sys.path.remove("/root/reference")
        for name in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[name]
    from gpbt_b200.state import EmulatorState
    stt = EmulatorState.from_trained(emu, keep_L=True)
    assert stt.kind == "Matern" and (stt.p, stt.n, stt.q, stt.m) == (3, 30, 3, 8)
    od = stt.oracle_dict()
    mean, cov = orc.emulator_predict(od, X, True, np.zeros(len(X)))
    assert np.max(np.abs(mean - ref["mean"]) / np.abs(ref["mean"])) <= 1e-9
    assert np.max(np.abs(cov - ref["cov"])) <= 1e-9 * np.max(np.abs(ref["cov"]))
    lp = orc.log_posterior([od], X, ref["lo"], ref["hi"], ref["y"], ref["c"])
    fin = np.isfinite(ref["lp"])
    assert np.array_equal(np.isfinite(lp), fin) and np.max(np.abs(lp[fin] - ref["lp"][fin])) <= 1e-8


def test_philox_known_answers():
    """Philox4x32-10 restatement (oracle/ensemble_oracle.py, mirrored by csrc/ensemble.cuh) against the
    known-answer vectors published with the generator (Random123 kat_vectors: counter, key -> output)."""
    from oracle.ensemble_oracle import philox4x32

    def run(ctr, key):
        return philox4x32(key[0] | (key[1] << 32), ctr[0] | (ctr[1] << 32), ctr[2], ctr[3])

    assert run([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_stretch_oracle_samples_a_known_gaussian():
    """The stretch-move restatement driven by the Philox streams samples a correlated 3-d Gaussian with
    the right mean and covariance (the published algorithm's invariant distribution)."""
    from oracle import ensemble_oracle as eo
    cov = np.array([[1.0, 0.6, 0.0], [0.6, 2.0, -0.5], [0.0, -0.5, 0.5]])
    mean = np.array([1.0, -2.0, 0.5])
    P = np.linalg.inv(cov)
    logp = lambda X: -0.5 * np.einsum("ni,ij,nj->n", X - mean, P, X - mean)
    nw, steps = 24, 1500
    u, partner, perm = eo.philox_streams(11, 0, steps, nw)
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01
    assert partner.min() == 0 and partner.max() == nw // 2 - 1
    x0 = mean + np.random.default_rng(0).normal(size=(nw, 3))
    chain, lps, acc = eo.stretch_run(logp, x0, logp(x0), u, partner, perm)
    flat = chain[300:].reshape(-1, 3)
    assert np.all(np.abs(flat.mean(axis=0) - mean) < 0.1)
    assert np.max(np.abs(np.cov(flat, rowvar=False) - cov)) < 0.15
    assert 0.3 < acc.mean() / steps < 0.8
    np.testing.assert_array_equal(lps[-1], logp(chain[-1]))
    assert sorted(eo.fixed_split(7)) == list(range(7)) and list(eo.fixed_split(5)) == [0, 2, 4, 1, 3]


def test_emulator_small_helpers_match_reference_formulas(tmp_path):
    """parametrization_* / _inverse_transform / outputPCAvsParam / sample_y of the drop-in Emulator
    (src/emulator.py:100-124, 243-248, 366-375, 608-633) -- CPU only, no device call."""
    from gpbt_b200 import synthetic
    from gpbt_b200.emulator import Emulator
    paths = synthetic.write_fixture(str(tmp_path), p=3, n=30, m=8)
    emu = Emulator(training_set_path=paths["train"], parameter_file=paths["par"], npc=3)
    # scalar formulas, written out independently
    assert emu.parametrization_zeta_over_s_vs_T(0.2, 0.18, 0.05, 0.02, 0.15, 0.3) == pytest.approx(
        0.2 * np.exp(-(0.15 - (0.18 - 0.15 * 0.09)) ** 2 / (2 * 0.02 ** 2)))
    assert emu.parametrization_zeta_over_s_vs_T(0.2, 0.18, 0.05, 0.02, 0.25, 0.0) == pytest.approx(
        0.2 * np.exp(-(0.25 - 0.18) ** 2 / (2 * 0.05 ** 2)))
    assert emu.parametrization_eta_over_s_vs_mu_B(0.1, 0.2, 0.4, 0.1) == pytest.approx(0.15)
    assert emu.parametrization_eta_over_s_vs_mu_B(0.1, 0.2, 0.4, 0.3) == pytest.approx(0.3)
    assert emu.parametrization_eta_over_s_vs_mu_B(0.1, 0.2, 0.4, 0.0) == pytest.approx(0.4)   # the reference's else branch
    assert emu.parametrization_y_loss_vs_y_init(1.0, 2.0, 4.0, 1.0) == pytest.approx(0.5)
    assert emu.parametrization_y_loss_vs_y_init(1.0, 2.0, 4.0, 3.0) == pytest.approx(1.5)
    assert emu.parametrization_y_loss_vs_y_init(1.0, 2.0, 4.0, 5.0) == pytest.approx(3.0)
    emu.trainEmulatorAutoMask()
    X, Z = emu.outputPCAvsParam()
    assert X.shape == (emu.nev, 3) and Z.shape == (3, emu.nev)
    back = emu._inverse_transform(Z.T)
    assert back.shape == (emu.nev, 8) and np.median(np.abs(back - emu.model_data) / np.abs(emu.model_data)) < 0.05
    ys = emu.sample_y(emu.design_points[:4], n_samples=5, random_state=0)
    assert ys.shape == (4, 5, 8) and np.all(np.isfinite(ys))


def test_emulator_band_host_side(tmp_path):
    """EmulatorBAND without a GPU: constructor on the reference's file formats, the surmise-free
    training error, state extraction from surmise-shaped fit information (hypind sharing, amplitudes)."""
    from gpbt_b200 import synthetic
    from gpbt_b200.emulator_band import EmulatorBAND
    from gpbt_b200.state import EmulatorState
    paths = synthetic.write_fixture(str(tmp_path), p=3, n=25, m=7)
    emu = EmulatorBAND(training_set_path=paths["train"], parameter_file=paths["par"], method="PCSK")
    assert emu.nobs == 7 and emu.nparameters == 3 and emu.design_min.shape == (3,)
    with pytest.raises(ValueError):
        EmulatorBAND(training_set_path=paths["train"], parameter_file=paths["par"], method="nope")
    with pytest.raises(ValueError):
        EmulatorBAND(training_set_path=paths["train"], parameter_file=paths["par"], exp_and_cov_diagonal=True)
    with pytest.raises(ImportError, match="surmise"):
        emu.trainEmulatorAutoMask()
    with pytest.raises(RuntimeError):
        emu.state
    info = synthetic.pcgp_fitinfo(3, 25, 7, 4)
    st = EmulatorState.from_pcgp_fitinfo(info, extravar_in_cov=True)
    assert st.kind == "PCGP" and (st.p, st.n, st.q, st.m) == (3, 25, 4, 7)
    g = np.exp(st.pcgp["hypcov"][:, -1])
    np.testing.assert_allclose(st.c + st.sn, 1.0 - st.pcgp["nug"])                 # (1-nug)(a + b), a + b = 1
    np.testing.assert_allclose(st.sn / st.c, g)
    np.testing.assert_array_equal(st.Linv, np.swapaxes(st.pcgp["Vh"], 1, 2))
    np.testing.assert_allclose(np.diag(st.Ctrunc), info["extravar"])
    np.testing.assert_array_equal(st.pcgp["hypcov"][1], st.pcgp["hypcov"][0])     # PC 1 shares PC 0's hyper-parameters
    # the fabricated fit interpolates its own training scores: r(theta_i) . pw = g_i up to the nugget
    from oracle import gp_oracle as orc
    zm, zv = orc.pc_predict(st.oracle_dict(), info["theta"])
    assert zv.max() < 0.05 * st.sig2.max() and np.all(zv >= 0)


def test_singular_truncation_plus_experiment_falls_back_to_dense():
    """F = blockdiag(Ctrunc) + cov_exp singular (zero truncation term, as on the PCGP path, and an
    experimental error of exactly 0 -- read_experiment_pickle turns NaN errors into 0): the chain must
    come up without low-rank factors (dense path), not raise at construction."""
    import warnings
    import gpbt_b200  # noqa: F401
    from types import SimpleNamespace
    from gpbt_b200.device import DeviceChain
    rng = np.random.default_rng(3)
    m, q = 6, 2
    st = SimpleNamespace(m=m, q=q, p_in=3, A=rng.normal(size=(q, m)), mu=np.zeros(m), Ctrunc=np.zeros((m, m)),
                         no_pca=False, exp_diag=False)
    cov_exp = np.diag([0.1, 0.2, 0.0, 0.3, 0.1, 0.2])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        dc = DeviceChain([st], np.zeros(3), np.ones(3), np.zeros(m), cov_exp)
    assert dc.lowrank is None and dc.lowrank_error and dc._checked
    assert any("dense path" in str(x.message) for x in w)


def test_default_devices_policy(monkeypatch):
    """single process: all visible devices, current first; one rank of a torchrun job: its own only;
    GPBT_DEVICES narrows or pins the list"""
    import gpbt_b200  # noqa: F401
    from gpbt_b200 import _lib, device
    n = _lib.lib.gpbt_device_count()
    monkeypatch.delenv("GPBT_DEVICES", raising=False)
    monkeypatch.setenv("WORLD_SIZE", "4")
    assert len(device.default_devices()) == 1
    monkeypatch.setenv("WORLD_SIZE", "1")
    monkeypatch.setenv("GPBT_DEVICES", "1")
    assert len(device.default_devices()) == 1
    monkeypatch.delenv("GPBT_DEVICES")
    assert len(device.default_devices()) == max(n, 1)
    monkeypatch.setenv("GPBT_DEVICES", "7,9")
    if n < 10:
        with pytest.raises(ValueError):
            device.default_devices()


def test_chain_state_key_sees_in_place_edits(tmp_path):
    """Chain.device() must notice in-place edits of the bounds / experimental data (the reference re-reads
    them on every call): the key is content based."""
    import pickle
    import gpbt_b200  # noqa: F401
    from gpbt_b200.mcmc import Chain
    syn = goldens.synthetic_module()
    paths = syn.write_fixture(str(tmp_path), 3, 12, 5)
    ch = Chain(mcmc_path=str(tmp_path / "mcmc" / "chain.pkl"), expdata_path=paths["exp"], model_parafile=paths["par"])
    ch.emuList = [object()]
    k0 = ch._state_key(False)
    ch.min[1] += 0.25
    k1 = ch._state_key(False)
    assert k1 != k0
    ch.expdata_cov[2, 2] *= 2.0                      # diagonal: seen by the per-call check
    k2 = ch._state_key(False)
    assert k2 != k1
    ch._dev_snapshot = ch.expdata_cov.copy()
    ch.expdata_cov[3, 1] += 1e-3                     # off the sampled set: seen by the full check
    assert ch._state_key(True) != ch._state_key(False) or ch._state_key(False) != k2
    ch.devices = [0]
    assert ch._state_key(False) != k2


def test_ptlmc_oracle_exchange_is_the_pinned_sweep(monkeypatch):
    """oracle/ptlmc_oracle.exchange_sweep (the checker of the device-resident PTLMC loop) = the sweep that is
    pinned to the unmodified reference (gpbt_b200.ptlmc.temp_exchange_python, tests/test_ptlmc.py), on the
    same slots and uniforms; and its Philox-keyed draws are slots in [1, n) / logs of uniforms"""
    from oracle import ptlmc_oracle as pto
    from gpbt_b200 import ptlmc
    rng = np.random.default_rng(0)
    for n in (2, 3, 37, 200):
        lp = 5.0 * rng.normal(size=n)
        temps = ptlmc.temperature_ladder(n - n // 4 - 1, n // 4 + 1, 25.0).ravel()
        slots, logu = pto.sweep_draws(11, 3, 1, n)
        assert slots.min() >= 1 and slots.max() <= n - 1 and np.all(logu <= 0.0)
        got = pto.exchange_sweep(lp, temps, np.arange(n), slots, logu)
        it = iter(logu)
        monkeypatch.setattr(np.random, "choice", lambda a, k: slots)
        monkeypatch.setattr(np.random, "uniform", lambda size=1: np.array([np.exp(next(it))]))
        want = ptlmc.temp_exchange_python(lp, temps, iters=1)
        monkeypatch.undo()
        assert np.array_equal(got, want)
    assert abs(pto.stride_of(-1.0) - 2 * (1 + np.tanh(-1.0))) < 1e-15
