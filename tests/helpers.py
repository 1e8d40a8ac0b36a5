"""Shared test helpers: build product states from golden files, error metrics, tolerances."""
import numpy as np

from tests import goldens

# tolerances from BASELINE.json north_star: 1e-9 relative on emulator means / variances (asserted on
# the observable-space outputs of the API boundary), 1e-8 absolute on log-likelihoods
REL = 1e-9
ABS_LP = 1e-8


# The north-star tolerances (1e-9 relative, 1e-8 absolute on log L) presuppose outputs that are determined
# to that accuracy.  The predictive variance is the difference k** - |L^-1 k|^2; in the reference-trained
# Matern emulator at the config-2 shape (c2_matern: the constant kernel ran into its upper bound 1e5) the
# two terms cancel by a factor of up to 7.9e6 (median 5e4; c2_rbf: 1.8e4 at most), so a last-bit change in
# the order of any sum -- sklearn's own `V.T @ V` against an einsum of the same numbers -- moves z_var by
# ~1e-8 relative and log L (|log L| ~ 400) by ~3e-6: the reference's golden values themselves are only
# defined to that accuracy (measured: the oracle, fed the reference's own L_, differs from them by 1.4e-8 /
# 3.1e-6).  PC-space variances are therefore checked everywhere with a backward-error term on the
# cancelling scale (var_tol), and c2_matern's observable-space outputs and log-likelihoods against the
# golden vectors with the tolerances below; against the oracle ON THE SAME STATE every case, c2_matern
# included, is held to what the arithmetic allows (see the tests).
ILL_CONDITIONED_TOL = {"c2_matern": (1e-6, 1e-4)}


def var_tol(ref_var, c, sn):
    """|error| allowed on PC-space variances [N, q]: 1e-9 relative plus 64 ulp of the cancelling terms c + sn"""
    return REL * np.abs(ref_var) + 64 * np.finfo(np.float64).eps * (np.asarray(c) + np.asarray(sn))[None, :]


def golden_tol(name):
    """(relative tolerance on means / variances, absolute tolerance on log-likelihoods) against the golden
    vectors of case `name`"""
    return ILL_CONDITIONED_TOL.get(name, (REL, ABS_LP))


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
