"""Shared test helpers: build product states from golden files, error metrics, tolerances."""
import numpy as np

from tests import goldens

# tolerances from BASELINE.json north_star: 1e-9 relative on emulator means / variances (asserted on
# the observable-space outputs of the API boundary), 1e-8 absolute on log-likelihoods
REL = 1e-9
ABS_LP = 1e-8


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def scaled_err(a, b):
    """max |a-b| / max |b|: for covariance matrices whose off-diagonal entries cancel to ~0"""
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def product_states(g):
    """EmulatorState list (product side) for golden dict g, plus the oracle dicts."""
    from gpbt_b200.state import EmulatorState
    sts = goldens.oracle_states(g)
    from gpbt_b200.state import ParamTrafo
    states = [EmulatorState.from_arrays(s["kind"], s["Xtr"], s["ell"], s["c"], s["sn"], s["alpha"],
                                        s["mu"], s["scale"], s.get("A"), s.get("Ctrunc"), L=s["L"],
                                        no_pca=s["no_pca"], exp_diag=s["exp_diag"],
                                        trafo=ParamTrafo(**s["trafo"]) if s.get("trafo") else None) for s in sts]
    return states, sts


def pc_scale(st, X):
    """sum_i |k_i alpha_i| per (row, PC): the natural scale for PC-space mean errors"""
    from oracle import gp_oracle as orc
    if st.get("trafo") is not None:
        X = orc.param_trafo(st["trafo"], X)
    out = np.empty((len(X), st["alpha"].shape[0]))
    for j in range(out.shape[1]):
        K = np.abs(orc.kernel_cross(X, st["Xtr"], st["c"][j], st["ell"][j], st["kind"]))
        out[:, j] = K @ np.abs(st["alpha"][j])
    return out
