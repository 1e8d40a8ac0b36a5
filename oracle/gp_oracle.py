"""
CPU oracle for the GPBayesTools-HIC hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (gpbayestools-hic_b200/) never imports it and has no CPU path.

It is a plain NumPy/SciPy restatement of what the reference computes between
`Chain.log_posterior(X)` and the LAPACK calls, following the reference operation by operation:

  reference (file:line, under /root/reference)          restated here
  ----------------------------------------------------  ---------------------------
  sklearn kernels.py RBF.__call__      (:1558-1571)      kernel_cross(kind="RBF")
  sklearn kernels.py Matern nu=1.5     (:1713-1729)      kernel_cross(kind="Matern")
  sklearn _gpr.py predict              (:446-475)        gp_predict_pc
  src/emulator.py Emulator.predict     (:465-605)        emulator_predict
  src/emulator.py parameterTrafoPCA    (:100-124,492-551) param_trafo
  src/emulator.py _inverse_transform   (:366-375)        emulator_predict (PCA branch)
  src/mcmc.py Chain._predict           (:153-166)        chain_predict
  src/mcmc.py mvn_loglike              (:23-65)          mvn_loglike
  src/mcmc.py Chain.log_likelihood     (:188-222)        log_likelihood
  src/mcmc.py Chain.log_posterior      (:261-299)        log_posterior

  src/emulator_BAND.py EmulatorBAND.predict (:386-478)  pcgp_pc_predict + emulator_predict (kind "PCGP")

Parity pin: tests/test_oracle_golden.py checks every function here against golden vectors that
tests/golden/make_golden.py produced by importing and running the UNMODIFIED reference
(/root/reference/src, scikit-learn 1.9.0) in the build container.

PARITY UNPINNED for the surmise (EmulatorBAND) path: `self.emu.predict(x, theta).mean() / .covx()` is
computed inside surmise 0.2.1 (requirements.txt:1), which is neither installed here nor vendored in
the reference tree, so no golden vector can be generated.  pcgp_covmat / pcgp_pc_predict restate the
published algorithm of surmise's emulationmethods/PCGP.py (`__covmat`, `predict`; PCSK's predict has
the same form) from its documentation and from SURVEY 8c; they pin the CUDA kernels to that
restatement, not to surmise itself.

An emulator "state" is a dict of float64 arrays (the trained quantities the reference keeps on
its sklearn objects):
    kind    "RBF" | "Matern"                       kernel family (src/emulator.py:288-300)
    Xtr     [n, p]   training design               gp.X_train_
    ell     [q, p]   anisotropic length scales     gp.kernel_.k1.k2.length_scale
    c       [q]      constant-kernel amplitude     gp.kernel_.k1.k1.constant_value
    sn      [q]      white-noise level             gp.kernel_.k2.noise_level
    alpha   [q, n]   K^-1 y                        gp.alpha_
    L       [q, n, n] lower Cholesky of K+0.1 I    gp.L_
    no_pca, exp_diag  bool                         perform_no_PCA_, exp_and_cov_diagonal_
    A       [q, m]   _trans_matrix[:npc]           (PCA mode)
    mu      [m]      scaler.mean_
    scale   [m]      scaler.scale_                 (no-PCA mode)
    Ctrunc  [m, m]   _cov_trunc                    (PCA mode)
"""
import numpy as np
from scipy.linalg import lapack, solve_triangular
from scipy.spatial.distance import cdist

# 2*log(0 + 1e-16) - 0/scale : the "extra_std prior" term with extra_std == 0
# (src/mcmc.py:199-222 and :281,296-297 -- extra_std is multiplied by 0.0 before use)
SYS_PRIOR_CONST = 2.0 * np.log(0.0 + 1e-16)


def kernel_cross(X, Xtr, c, ell, kind):
    """c * kappa(X/ell, Xtr/ell); the WhiteKernel contributes zero to cross terms
    (sklearn kernels.py:1406-1419).  RBF: kernels.py:1569-1570.  Matern nu=1.5: :1722-1727."""
    d = cdist(X / ell, Xtr / ell, metric="sqeuclidean" if kind == "RBF" else "euclidean")
    if kind == "RBF":
        return c * np.exp(-0.5 * d)
    if kind == "Matern":
        r = d * np.sqrt(3.0)
        return c * ((1.0 + r) * np.exp(-r))
    raise ValueError(kind)


def gp_predict_pc(state, j, X):
    """Mean and the DIAGONAL of the predictive covariance of GP j (sklearn _gpr.py:446-466).
    diag(kernel_(X)) = c + sn (Constant*RBF/Matern diag is c; White adds sn), no clamping."""
    K = kernel_cross(X, state["Xtr"], state["c"][j], state["ell"][j], state["kind"])
    mean = K @ state["alpha"][j]
    V = solve_triangular(state["L"][j], K.T, lower=True, check_finite=False)
    var = (state["c"][j] + state["sn"][j]) - np.einsum("ij,ij->j", V, V)
    return mean, var


def _zeta_over_s(zeta_max, T_zeta0, sigma_plus, sigma_minus, T, mu_B=0.0):
    """src/emulator.py:100-106"""
    T_zeta_muB = T_zeta0 - 0.15 * mu_B ** 2.
    sig = np.where(T < T_zeta0, sigma_minus, sigma_plus)
    return zeta_max * np.exp(-(T - T_zeta_muB) ** 2. / (2. * sig ** 2.))


def _eta_over_s(eta_0, eta_2, eta_4, mu_B):
    """src/emulator.py:109-115"""
    return np.where((0. < mu_B) & (mu_B <= 0.2), eta_0 + (eta_2 - eta_0) * (mu_B / 0.2),
                    np.where((0.2 < mu_B) & (mu_B < 0.4), eta_2 + (eta_4 - eta_2) * ((mu_B - 0.2) / 0.2), eta_4))


def _y_loss(yloss_2, yloss_4, yloss_6, y_init):
    """src/emulator.py:118-124"""
    return np.where((0. < y_init) & (y_init <= 2.), yloss_2 * (y_init / 2.),
                    np.where((2. < y_init) & (y_init < 4.), yloss_2 + (yloss_4 - yloss_2) * ((y_init - 2.) / 2.),
                             yloss_4 + (yloss_6 - yloss_4) * ((y_init - 4.) / 2.)))


_CURVES = {0: _zeta_over_s, 1: _eta_over_s, 2: _y_loss}


def param_trafo(trafo, X):
    """The parameterTrafoPCA pre-transform inside Emulator.predict (src/emulator.py:492-551):
    curves on their grids -> StandardScaler.transform -> PCA.transform; untouched columns first, then
    the bulk, shear and y_loss components."""
    X = np.asarray(X, dtype=np.float64)
    used = {int(c) for g in trafo["groups"] for c in g["idx"]}
    cols = [X[:, [c for c in range(trafo["p_in"]) if c not in used]]]
    for g in trafo["groups"]:
        grid = np.linspace(*g["grid"])
        args = [X[:, int(c)][:, None] for c in g["idx"]]
        f = _CURVES[int(g["kind"])](*args, grid[None, :])
        scaled = (f - g["smean"]) / g["sscale"]
        cols.append((scaled - g["pmean"]) @ g["comp"].T)
    return np.concatenate(cols, axis=1)


def pcgp_covmat(x1, x2, gammav):
    """surmise PCGP `__covmat`: separable Matern-3/2 in |x1_d - x2_d| / exp(gamma_d) mixed with a
    constant, weights 1/(1+e^g) and e^g/(1+e^g) for g = gammav[-1].  [UNPINNED restatement]"""
    x1, x2 = np.atleast_2d(x1), np.atleast_2d(x2)
    V = np.zeros((x1.shape[0], x2.shape[0]))
    R = np.full((x1.shape[0], x2.shape[0]), 1.0 / (1.0 + np.exp(gammav[-1])))
    for k in range(len(gammav) - 1):
        S = np.abs(np.subtract.outer(x1[:, k], x2[:, k]) / np.exp(gammav[k]))
        R *= (1.0 + S)
        V -= S
    R *= np.exp(V)
    R += np.exp(gammav[-1]) / (1.0 + np.exp(gammav[-1]))
    return R


def pcgp_pc_predict(state, theta):
    """surmise PCGP `predict`, PC space: predvec_k = r pw_k, predvar_k = sig2_k |1 - |r Vh_k|^2| with
    r = (1 - nug_k) covmat(theta, theta_train, hypcov_k).  [UNPINNED restatement]"""
    q = state["pw"].shape[0]
    zm, zv = np.empty((theta.shape[0], q)), np.empty((theta.shape[0], q))
    for k in range(q):
        r = (1.0 - state["nug"][k]) * pcgp_covmat(theta, state["Xtr"], state["hypcov"][k])
        rVh = r @ state["Vh"][k]
        zm[:, k] = r @ state["pw"][k]
        zv[:, k] = state["sig2"][k] * np.abs(1.0 - np.sum(rVh ** 2, axis=1))
    return zm, zv


def pc_predict(state, X, extra_std=None):
    """z_mean[N,q], z_var[N,q] for all GPs; z_var includes extra_std**2 (src/emulator.py:573-579).
    For a PCGP state extra_std is ignored, as EmulatorBAND.predict ignores it (src/emulator_BAND.py:386)."""
    X = np.asarray(X, dtype=np.float64)
    if state.get("trafo") is not None:
        X = param_trafo(state["trafo"], X)
    if state["kind"] == "PCGP":
        return pcgp_pc_predict(state, X)
    q = state["alpha"].shape[0]
    zm = np.empty((X.shape[0], q))
    zv = np.empty((X.shape[0], q))
    for j in range(q):
        zm[:, j], zv[:, j] = gp_predict_pc(state, j, X)
    if extra_std is not None:
        zv += np.asarray(extra_std, dtype=np.float64).reshape(-1, 1) ** 2
    return zm, zv


def emulator_predict(state, X, return_cov=True, extra_std=None):
    """Emulator.predict (src/emulator.py:465-605); the parameterTrafoPCA pre-transform is applied in
    pc_predict when the state carries one."""
    X = np.asarray(X, dtype=np.float64)
    zm, zv = pc_predict(state, X, extra_std)
    if not state["no_pca"]:
        mean = zm @ state["A"] + state["mu"]                         # :366-375
    else:
        mean = zm * state["scale"] + state["mu"]                     # StandardScaler.inverse_transform, :562-565
    if state["exp_diag"]:
        mean = np.exp(mean)                                          # :567-568
    if not return_cov:
        return mean
    N, m = mean.shape
    if not state["no_pca"]:
        A = state["A"]
        var_trans = np.einsum("ki,kj->kij", A, A).reshape(A.shape[0], m * m)   # :353-355
        cov = (zv @ var_trans).reshape(N, m, m) + state["Ctrunc"]               # :584-587
    else:
        cov = np.zeros((N, m, m))
        idx = np.arange(m)
        cov[:, idx, idx] = zv                                                  # :588-592
    if state["exp_diag"]:
        idx = np.arange(m)
        fstd = np.sqrt(cov[:, idx, idx])
        cov = np.zeros((N, m, m))
        cov[:, idx, idx] = (fstd * mean) ** 2                                  # :594-601
    return mean, cov


def chain_predict(states, X, extra_std=0.0):
    """Chain._predict (src/mcmc.py:153-166): concatenated means, block-diagonal covariance."""
    X = np.asarray(X, dtype=np.float64)
    N = X.shape[0]
    extra = extra_std * X[:, -1]
    means, covs = [], []
    for st in states:
        mu, cv = emulator_predict(st, X, True, extra)
        means.append(mu)
        covs.append(cv)
    m = sum(x.shape[1] for x in means)
    mean = np.concatenate(means, axis=1)
    cov = np.zeros((N, m, m))
    o = 0
    for cv in covs:
        k = cv.shape[1]
        cov[:, o:o + k, o:o + k] = cv
        o += k
    return mean, cov


def mvn_loglike(y, cov):
    """src/mcmc.py:23-65: dpotrf (upper, default) + dpotrs; unnormalised log density."""
    U, info = lapack.dpotrf(cov, clean=False)
    a, info2 = lapack.dpotrs(U, y)
    return -0.5 * np.dot(y, a) - np.log(U.diagonal()).sum()


def _log_target(states, X, lo, hi, y_exp, cov_exp, oob_value):
    X = np.array(X, dtype=np.float64, ndmin=2)
    lp = np.zeros(X.shape[0])
    inside = np.all((X > lo) & (X < hi), axis=1)
    lp[~inside] = oob_value
    if np.count_nonzero(inside) > 0:
        mean, cov = chain_predict(states, X[inside], 0.0)
        dY = mean - y_exp
        cov = cov + cov_exp
        lp[inside] += np.array([mvn_loglike(d, c) for d, c in zip(dY, cov)])
        lp[inside] += SYS_PRIOR_CONST
    return lp


def log_posterior(states, X, lo, hi, y_exp, cov_exp):
    """Chain.log_posterior (src/mcmc.py:261-299); no log-prior term is added by the reference."""
    return _log_target(states, X, lo, hi, y_exp, cov_exp, -np.inf)


def log_likelihood(states, X, lo, hi, y_exp, cov_exp, finite=False):
    """Chain.log_likelihood (src/mcmc.py:188-222)."""
    return _log_target(states, X, lo, hi, y_exp, cov_exp, -1e300 if finite else -np.inf)


# ---------------------------------------------------------------------------------------------
# chunked driver used only by bench.py's CPU-baseline legs (the restatement above is O(N) per
# row, unlike the reference whose sklearn call builds an N x N matrix)
# ---------------------------------------------------------------------------------------------
def log_posterior_chunked(states, X, lo, hi, y_exp, cov_exp, chunk=256):
    out = np.empty(len(X))
    for s in range(0, len(X), chunk):
        out[s:s + chunk] = log_posterior(states, X[s:s + chunk], lo, hi, y_exp, cov_exp)
    return out
