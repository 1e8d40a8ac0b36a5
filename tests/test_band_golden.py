"""EmulatorBAND / surmise path against golden vectors of the real library.  The vectors do not exist yet:
surmise 0.2.1 is absent from the build image (tests/golden/make_golden_band.py generates them the moment it
is importable).  Until then these tests xfail with that reason, so the "parity unpinned" row flips by
running one script and committing two files."""
import os

import numpy as np
import pytest

from oracle import gp_oracle as orc
from tests import goldens
from tests.helpers import ABS_LP, REL, rel_err, scaled_err

METHODS = ["pcgp", "pcsk"]


def _load(method):
    path = os.path.join(goldens.GOLDEN_DIR, "band_%s.npz" % method)
    if not os.path.exists(path):
        pytest.xfail("surmise absent: tests/golden/band_%s.npz has not been generated "
                     "(run tests/golden/make_golden_band.py where surmise==0.2.1 is installed)" % method)
    return goldens.load("band_%s" % method)


def _fitinfo(g):
    q = int(g["n_pc"])
    emul = [dict(hypcov=g["pc%d_hypcov" % k], hypind=int(g["pc%d_hypind" % k]), nug=float(g["pc%d_nug" % k]),
                 Vh=g["pc%d_Vh" % k], pw=g["pc%d_pw" % k], sig2=float(g["pc%d_sig2" % k])) for k in range(q)]
    return {"theta": g["theta"], str(g["pct_key"]): g["pct"], "scale": g["scale"], "offset": g["offset"],
            "extravar": g["extravar"], "emulist": emul}


def _state(g):
    """the emulator state whose covariance convention matches the library's own covx() output"""
    import gpbt_b200  # noqa: F401
    from gpbt_b200.state import EmulatorState
    info = _fitinfo(g)
    best = None
    for extravar_in_cov in (False, True):
        st = EmulatorState.from_pcgp_fitinfo(info, extravar_in_cov=extravar_in_cov)
        _, cov = orc.emulator_predict(st.oracle_dict(), g["X"][g["inside"]][:4], True, None)
        err = scaled_err(cov, g["cov"][:4])
        if best is None or err < best[0]:
            best = (err, extravar_in_cov, st)
    return best


@pytest.mark.parametrize("method", METHODS)
def test_oracle_restatement_matches_surmise(method):
    g = _load(method)
    err, extravar_in_cov, st = _state(g)
    Xi = g["X"][g["inside"]]
    mean, cov = orc.emulator_predict(st.oracle_dict(), Xi, True, None)
    assert rel_err(mean, g["mean"]) <= REL and scaled_err(cov, g["cov"]) <= REL, (method, extravar_in_cov)
    lp = orc.log_posterior([st.oracle_dict()], g["X"], g["lo"], g["hi"], g["y_exp"], g["cov_exp"])
    fin = np.isfinite(g["lp_posterior"])
    assert np.array_equal(np.isfinite(lp), fin) and np.max(np.abs(lp[fin] - g["lp_posterior"][fin])) <= ABS_LP


@pytest.mark.gpu
@pytest.mark.parametrize("method", METHODS)
def test_cuda_path_matches_surmise(method):
    from gpbt_b200.device import DeviceChain, DeviceEmulator
    g = _load(method)
    _, extravar_in_cov, st = _state(g)
    Xi = g["X"][g["inside"]]
    mean, cov = DeviceEmulator(st).predict(Xi, return_cov=True)
    assert rel_err(mean, g["mean"]) <= REL and scaled_err(cov, g["cov"]) <= REL, (method, extravar_in_cov)
    ch = DeviceChain([st], g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
    fin = np.isfinite(g["lp_posterior"])
    for path in ("auto", "dense"):
        lp = ch.log_target(g["X"], -np.inf, path=path)
        assert np.array_equal(np.isfinite(lp), fin) and np.max(np.abs(lp[fin] - g["lp_posterior"][fin])) <= ABS_LP
    lf = ch.log_target(g["X"], -1e300)
    assert np.max(np.abs(lf - g["lp_like_finite"])) <= ABS_LP
    ch.release()
