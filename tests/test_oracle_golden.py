"""Pin the oracle (oracle/gp_oracle.py) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Tolerances are those BASELINE.json's north_star states: 1e-9
relative on emulator means/variances (asserted on the observable-space outputs),
1e-8 absolute on log-likelihoods."""
import numpy as np
import pytest

from oracle import gp_oracle as orc
from tests import goldens
from tests.helpers import golden_tol, var_tol

CASES = [c for c in goldens.SMALL_CASES + ["c2_rbf", "c2_matern"] if c in goldens.available()]


def rel_err(a, b, floor=0.0):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor + 1e-300))


@pytest.mark.parametrize("case", CASES)
def test_pc_space(case):
    REL, ABS_LP = golden_tol(case)
    g = goldens.load(case)
    Xin = g["X"][g["inside"]]
    for e, st in enumerate(goldens.oracle_states(g)):
        if "e%d_z_mean" % e not in g:
            continue   # parameterTrafoPCA case: pinned through the observable-space outputs only
        zm, zv = orc.pc_predict(st, Xin)
        # PC-space means are ill-conditioned sums: compare against the scale of the
        # summands, not of the (possibly cancelling) result
        K = np.abs(orc.kernel_cross(Xin, st["Xtr"], st["c"][0], st["ell"][0], st["kind"]))
        scale = (K @ np.abs(st["alpha"][0])).max()
        assert np.max(np.abs(zm - g["e%d_z_mean" % e])) <= 1e-12 * max(scale, 1.0) * 10
        assert np.all(np.abs(zv - g["e%d_z_var" % e]) <= var_tol(g["e%d_z_var" % e], st["c"], st["sn"]))


@pytest.mark.parametrize("case", CASES)
def test_emulator_predict(case):
    REL, ABS_LP = golden_tol(case)
    g = goldens.load(case)
    Xin = g["X"][g["inside"]]
    for e, st in enumerate(goldens.oracle_states(g)):
        rows = g["e%d_mean_x" % e].shape[0]
        mean, cov = orc.emulator_predict(st, Xin[:rows], True, g["extra_std"][:rows])
        assert rel_err(mean, g["e%d_mean_x" % e]) <= REL
        ref = g["e%d_cov_x" % e]
        assert np.max(np.abs(cov - ref)) <= REL * np.max(np.abs(ref))
        d = np.arange(ref.shape[1])
        assert rel_err(cov[:, d, d], ref[:, d, d]) <= REL
        mean0 = orc.emulator_predict(st, Xin[:256], False)
        assert rel_err(mean0, g["e%d_mean0" % e]) <= REL


@pytest.mark.parametrize("case", CASES)
def test_chain_and_loglike(case):
    REL, ABS_LP = golden_tol(case)
    g = goldens.load(case)
    states = goldens.oracle_states(g)
    Xin = g["X"][g["inside"]]
    rows = g["chain_mean"].shape[0]
    mean, cov = orc.chain_predict(states, Xin[:rows], 0.0)
    assert rel_err(mean, g["chain_mean"]) <= REL
    assert np.max(np.abs(cov - g["chain_cov"])) <= REL * np.max(np.abs(g["chain_cov"]))
    mvn_y, mvn_cov = g["chain_mean"] - g["y_exp"], g["chain_cov"] + g["cov_exp"]
    vals = np.array([orc.mvn_loglike(y, c) for y, c in zip(mvn_y, mvn_cov)])
    assert np.max(np.abs(vals - g["mvn_val"])) <= ABS_LP
    args = (states, g["X"], g["lo"], g["hi"], g["y_exp"], g["cov_exp"])
    lp = orc.log_posterior(*args)
    ref = g["lp_posterior"]
    assert np.array_equal(np.isneginf(lp), np.isneginf(ref))
    fin = np.isfinite(ref)
    assert np.max(np.abs(lp[fin] - ref[fin])) <= ABS_LP
    lf = orc.log_likelihood(*args, finite=True)
    assert np.array_equal(lf == -1e300, g["lp_like_finite"] == -1e300)
    assert np.max(np.abs(lf[fin] - g["lp_like_finite"][fin])) <= ABS_LP
    ls = orc.log_posterior(states, g["X"], g["lo"], g["hi"], g["y_exp"], goldens.cov_exp_sys(g))
    assert np.max(np.abs(ls[fin] - g["lp_posterior_sys"][fin])) <= ABS_LP
