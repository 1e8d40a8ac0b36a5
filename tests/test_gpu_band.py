"""GPU tests of the surmise PCGP / PCSK path (EmulatorBAND.predict, src/emulator_BAND.py:386-478):
kernel (a) KIND 2 + kernel (b) + the log-likelihood paths against the oracle's restatement of
surmise's published predict.  PARITY UNPINNED against surmise itself (not installed, not vendored):
what these tests pin is the CUDA path to that restatement, on synthetic surmise-shaped fits."""
import numpy as np
import pytest

from oracle import gp_oracle as orc
from tests.helpers import ABS_LP, REL, rel_err, scaled_err

pytestmark = pytest.mark.gpu


def fit(p, n, m, q, **kw):
    from gpbt_b200 import synthetic
    from gpbt_b200.state import EmulatorState
    info = synthetic.pcgp_fitinfo(p, n, m, q, **kw)
    lo, hi = synthetic.box(p)
    return info, EmulatorState.from_pcgp_fitinfo(info), lo, hi


def mean_scale(od, X):
    """sum_i |r_i pw_i| per (row, PC): the natural scale of the PC-space mean's rounding error"""
    out = np.empty((len(X), od["pw"].shape[0]))
    for k in range(out.shape[1]):
        r = (1 - od["nug"][k]) * orc.pcgp_covmat(X, od["Xtr"], od["hypcov"][k])
        out[:, k] = np.abs(r) @ np.abs(od["pw"][k])
    return out


@pytest.mark.parametrize("p,n,m,q", [(5, 80, 30, 6), (7, 100, 50, 10), (15, 260, 40, 5), (2, 33, 9, 3)])
def test_pc_space_against_restatement(p, n, m, q):
    import torch
    from gpbt_b200.device import DeviceEmulator
    info, st, lo, hi = fit(p, n, m, q)
    od = st.oracle_dict()
    rng = np.random.default_rng(p)
    X = np.concatenate((rng.uniform(lo, hi, (301, p)), info["theta"][:7]))      # incl. design points (variance -> small)
    de = DeviceEmulator(st)
    zm, zv = de.pc_predict_device(torch.from_numpy(X).cuda())
    zm, zv = zm.cpu().numpy(), zv.cpu().numpy()
    om, ov = orc.pc_predict(od, X)
    assert np.max(np.abs(zm - om) / mean_scale(od, X)) <= 1e-12
    # the variance is sig2 |1 - |r Vh|^2|: error measured against sig2 (the cancellation is the algorithm's)
    assert np.max(np.abs(zv - ov) / st.sig2) <= 1e-11
    assert np.all(zv >= 0)
    # extra_std is accepted and ignored
    zm2, zv2 = de.pc_predict_device(torch.from_numpy(X).cuda(), torch.full((len(X),), 0.3, dtype=torch.float64).cuda())
    assert torch.equal(zv2.cpu(), torch.from_numpy(zv))


@pytest.mark.parametrize("exp_diag", [False, True])
def test_emulator_band_predict(exp_diag):
    from gpbt_b200.emulator_band import EmulatorBAND
    info, st, lo, hi = fit(6, 90, 24, 8)
    if exp_diag:   # log-observable emulator: keep exp() tame
        info = dict(info, offset=info["offset"] * 0.1, scale=info["scale"] * 0.1)
    emu = EmulatorBAND.from_fitinfo(info, lo, hi, exp_and_cov_diagonal=exp_diag)
    X = np.random.default_rng(1).uniform(lo, hi, (57, 6))
    mean, cov = emu.predict(X, return_cov=True, extra_std=0.7)
    od = emu.state.oracle_dict()
    omean, ocov = orc.emulator_predict(od, X, True, None)
    assert mean.shape == (57, 24) and cov.shape == (57, 24, 24)
    assert rel_err(mean, omean) <= REL and scaled_err(cov, ocov) <= REL
    assert rel_err(emu.predict(X, return_cov=False), omean) <= REL
    m2, var = emu.predict_diag(X)
    assert rel_err(var, np.array([c.diagonal() for c in ocov])) <= 1e-8
    assert emu.predict(X[0]).__class__ is tuple and emu.predict(X[0])[0].shape == (1, 24)   # 1-D input -> one row


@pytest.mark.parametrize("path", ["lowrank", "dense"])
def test_chain_on_band_emulators(path):
    """Chain.log_posterior over a PCGP emulator and a sklearn-GP emulator together."""
    from gpbt_b200.device import DeviceChain
    from gpbt_b200.state import EmulatorState
    from gpbt_b200 import synthetic
    info, st, lo, hi = fit(5, 70, 20, 6)
    info2, st2, _, _ = fit(5, 64, 12, 4, seed=5, shared_hyper=False)
    st2 = EmulatorState.from_pcgp_fitinfo(info2, extravar_in_cov=True)
    arr = synthetic.untrained_state_arrays(5, 50, 16, 5)
    st3 = EmulatorState.from_arrays(**arr, keep_L=True)
    states = [st, st2, st3]
    ods = [s.oracle_dict() for s in states]
    M = 20 + 12 + 16
    rng = np.random.default_rng(3)
    X = rng.uniform(lo - 0.02 * (hi - lo), hi + 0.02 * (hi - lo), (400, 5))
    y = np.concatenate([orc.emulator_predict(o, (0.45 * (lo + hi))[None, :], False)[0] for o in ods])
    B = rng.normal(size=(M, 3)) * 0.05
    cov_exp = np.diag((0.03 * np.abs(y)) ** 2 + 1e-4) + B @ B.T      # couples the emulators
    ch = DeviceChain(states, lo, hi, y, cov_exp)
    lp = ch.log_target(X, -np.inf, path=path)
    want = orc.log_posterior(ods, X, lo, hi, y, cov_exp)
    fin = np.isfinite(want)
    assert 0 < fin.sum() < len(X) and np.array_equal(np.isneginf(lp), ~fin)
    assert np.max(np.abs(lp[fin] - want[fin])) <= ABS_LP
    if path == "lowrank":
        assert ch.lowrank is not None and ch.lowrank_check["max_abs_diff"] <= 1e-8
    ch.release()


def test_config3_shape():
    """BASELINE config 3: 15 parameters, 1000 design points (dense 1000 x 1000 Vh per PC, L2-resident
    no more), a batch of PTLMC-sized calls."""
    from gpbt_b200.device import DeviceChain
    info, st, lo, hi = fit(15, 1000, 60, 4)
    od = st.oracle_dict()
    rng = np.random.default_rng(8)
    X = rng.uniform(lo, hi, (96, 15))
    y = orc.emulator_predict(od, (0.5 * (lo + hi))[None, :], False)[0]
    cov_exp = np.diag((0.03 * np.abs(y)) ** 2)
    ch = DeviceChain([st], lo, hi, y, cov_exp)
    lp = ch.log_target(X, -np.inf)
    want = orc.log_posterior([od], X, lo, hi, y, cov_exp)
    assert np.all(np.isfinite(want)) and np.max(np.abs(lp - want)) <= ABS_LP
    one = ch.log_target(X[5], -np.inf)
    assert one.shape == (1,) and abs(one[0] - lp[5]) <= 1e-10
    ch.release()


def test_from_trained_accepts_a_band_like_object():
    """Chain.loadEmulator hands dill-loaded objects to EmulatorState.from_trained: an EmulatorBAND-shaped
    object (attribute `emu` holding the surmise emulator with its `_info`) is recognised."""
    from gpbt_b200 import synthetic
    from gpbt_b200.state import EmulatorState
    info = synthetic.pcgp_fitinfo(4, 40, 10, 3)
    surmise_like = type("emulator", (), {"_info": info})()
    band_like = type("EmulatorBAND", (), {"emu": surmise_like, "exp_and_cov_diagonal_": False,
                                          "parameterTrafoPCA_": False})()
    st = EmulatorState.from_trained(band_like)
    assert st.kind == "PCGP" and st.q == 3 and st.n == 40 and st.m == 10
    assert np.array_equal(st.pcgp["hypcov"][1], st.pcgp["hypcov"][0])      # hypind sharing expanded
    assert st.handle() is not None
