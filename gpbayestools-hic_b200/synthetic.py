"""
Deterministic synthetic designs in the reference's on-disk formats.

The reference ships no data (examples/EmulatorTraining.ipynb reads ../data/*.pkl which are not
in the repository), so every test / bench fixture is produced here:

  * training pickle  {str(id): {"parameter": [p], "obs": [2, m]}}   (src/emulator.py:378-415)
  * experiment pickle, same layout with one entry                    (src/mcmc.py:302-324)
  * parameter file   "name: label, min, max"                         (src/__init__.py:21-32)

The recipe (box, Latin-hypercube design, tanh-network simulator) is the one SURVEY.md 8d states.
"""
import os
import pickle

import numpy as np
from scipy.stats import qmc

# BASELINE.json configs: (p params, n design points, m observables, q principal components)
SHAPES = {
    "C1": dict(p=5, n=100, m=50, q=10),
    "C2": dict(p=17, n=500, m=300, q=20),
    "C3": dict(p=15, n=1000, m=300, q=20),
}


def box(p):
    lo = np.zeros(p)
    hi = 1.0 + 0.25 * np.arange(p)
    return lo, hi


class Simulator:
    """Y = 5 + tanh(2 U W1) W2 + 0.5 sin(3 U_0) linspace(0,1,m),  U = (x-lo)/(hi-lo)."""

    def __init__(self, p, m, seed=20261018, hidden=24):
        rng = np.random.default_rng(seed)
        self.p, self.m = p, m
        self.lo, self.hi = box(p)
        self.W1 = rng.normal(0.0, np.sqrt(1.0 / p), (p, hidden))
        self.W2 = rng.normal(0.0, np.sqrt(1.0 / hidden), (hidden, m))
        self.ramp = np.linspace(0.0, 1.0, m)

    def unit(self, X):
        return (np.asarray(X, dtype=np.float64) - self.lo) / (self.hi - self.lo)

    def __call__(self, X):
        U = np.atleast_2d(self.unit(X))
        H = np.tanh(2.0 * U @ self.W1)
        return 5.0 + H @ self.W2 + 0.5 * np.sin(3.0 * U[:, :1]) * self.ramp


def design(p, n, seed=7):
    lo, hi = box(p)
    return lo + (hi - lo) * qmc.LatinHypercube(d=p, seed=seed).random(n)


def training_dict(p, n, m, noise=0.01, seed=11):
    sim = Simulator(p, m)
    Xtr = design(p, n)
    rng = np.random.default_rng(seed)
    Y = sim(Xtr) + noise * rng.standard_normal((n, m))
    return {str(i): {"parameter": Xtr[i].copy(),
                     "obs": np.stack([Y[i], np.full(m, noise)])} for i in range(n)}


def experiment_dict(p, m, rel_err=0.03, u0=0.4):
    sim = Simulator(p, m)
    x0 = sim.lo + u0 * (sim.hi - sim.lo)
    y = sim(x0)[0]
    return {"0": {"parameter": x0, "obs": np.stack([y, rel_err * np.abs(y)])}}


def parameter_file_text(p):
    lo, hi = box(p)
    lines = ["# synthetic design box", "# format: parameter_name: label, min, max"]
    for d in range(p):
        lines.append("par%d: $\\theta_{%d}$, %r, %r" % (d, d, float(lo[d]), float(hi[d])))
    return "\n".join(lines) + "\n"


def write_fixture(outdir, p, n, m, tag=""):
    """Write train/exp pickles + parameter file; returns their paths."""
    os.makedirs(outdir, exist_ok=True)
    paths = dict(train=os.path.join(outdir, "train%s.pkl" % tag),
                 exp=os.path.join(outdir, "exp%s.pkl" % tag),
                 par=os.path.join(outdir, "par%s.txt" % tag))
    with open(paths["train"], "wb") as f:
        pickle.dump(training_dict(p, n, m), f)
    with open(paths["exp"], "wb") as f:
        pickle.dump(experiment_dict(p, m), f)
    with open(paths["par"], "w") as f:
        f.write(parameter_file_text(p))
    return paths


def walkers(p, N, seed=3, frac_outside=0.01):
    """Uniform walkers in the box with a fixed fraction of rows pushed out of bounds."""
    lo, hi = box(p)
    rng = np.random.default_rng(seed)
    X = rng.uniform(lo, hi, (N, p))
    n_out = int(round(frac_outside * N))
    if n_out:
        rows = rng.choice(N, n_out, replace=False)
        cols = rng.integers(0, p, n_out)
        X[rows, cols] = hi[cols] + 0.1
    return X


def systematic_cov(m, rank=5, scale=0.02, seed=5):
    """Random PSD low-rank 'systematic' term for the full-covariance sweep (BASELINE config 4)."""
    rng = np.random.default_rng(seed)
    G = scale * rng.standard_normal((m, rank))
    return G @ G.T


def untrained_state_arrays(p, n, m, q, kind="RBF", seed=17, ell_range=(0.8, 4.0), c=50.0, sn=0.01):
    """A self-consistent emulator state of a given shape WITHOUT the L-BFGS hyper-parameter search
    (which takes minutes at n = 1000): data from the synthetic simulator, StandardScaler + whitened
    PCA as src/emulator.py:260-272 does, fixed plausible kernel hyper-parameters, and alpha_ / L_
    exactly as sklearn's fit builds them from those (_gpr.py:349-367).  For throughput runs at
    shapes that have no golden file; keyword arguments for EmulatorState.from_arrays."""
    from scipy.linalg import cho_solve
    from .state import cholesky_of_kernel
    rng = np.random.default_rng(seed)
    sim = Simulator(p, m)
    Xtr = design(p, n)
    Y = sim(Xtr) + 0.01 * rng.standard_normal((n, m))
    mu, scale = Y.mean(0), Y.std(0)
    Ys = (Y - mu) / scale
    U, S, Vt = np.linalg.svd(Ys, full_matrices=False)
    expl_var = S ** 2 / (n - 1)
    Z = (U * S)[:, :q] / np.sqrt(expl_var[:q])                      # whitened PC scores
    trans = Vt * np.sqrt(expl_var)[:, None] * scale                 # _trans_matrix (src/emulator.py:335-339)
    A, B = trans[:q], trans[q:]
    Ctrunc = B.T @ B
    Ctrunc[np.diag_indices(m)] += 1e-4 * scale ** 2
    span = box(p)[1] - box(p)[0]
    ell = span * rng.uniform(ell_range[0], ell_range[1], (q, p))
    cs, sns = np.full(q, c), np.full(q, sn)
    L = np.stack([cholesky_of_kernel(kind, Xtr, cs[j], ell[j], sns[j]) for j in range(q)])
    alpha = np.stack([cho_solve((L[j], True), Z[:, j]) for j in range(q)])
    return dict(kind=kind, Xtr=Xtr, ell=ell, c=cs, sn=sns, alpha=alpha, mu=mu, scale=scale, A=A,
                Ctrunc=Ctrunc, L=L)


def pcgp_fitinfo(p, n, m, q, seed=23, nug=1e-4, shared_hyper=True):
    """A surmise-PCGP-shaped fit (`emulator._info`) built with NumPy on the synthetic simulator, for
    tests and benchmarks of the EmulatorBAND path -- surmise itself is not installed here, so the
    hyper-parameters are drawn, not optimised: standardised outputs -> SVD principal components ->
    per PC the separable-Matern correlation R, its eigen-factor Vh = V / sqrt(W), pw = R^-1 g and
    sig2 = g . pw / n.  Pairs of PCs share hyper-parameters (hypind) as surmise's fit does."""
    rng = np.random.default_rng(seed)
    lo, hi = box(p)
    theta = design(p, n, seed=seed)
    W1 = rng.normal(0, 1 / np.sqrt(p), (p, 24))
    W2 = rng.normal(0, 1 / np.sqrt(24), (24, m))
    U = (theta - lo) / (hi - lo)
    F = 5 + np.tanh(2 * U @ W1) @ W2 + 0.5 * np.sin(3 * U[:, :1]) * np.linspace(0, 1, m) + 0.01 * rng.normal(size=(n, m))
    offset, scale = F.mean(0), F.std(0)
    Z = (F - offset) / scale
    _, S, Vt = np.linalg.svd(Z, full_matrices=False)
    pct = (Vt[:q].T * S[:q] / np.sqrt(n))                       # [m, q]
    G = Z @ np.linalg.pinv(pct).T                               # PC scores [n, q]
    emulist = []
    for k in range(q):
        hypind = k - (k % 2) if shared_hyper else k
        if hypind == k:
            hyp = np.concatenate((np.log((hi - lo) * rng.uniform(0.6, 2.5, p)), [rng.uniform(-4.0, -2.0)]))
        else:
            hyp = emulist[hypind]["hypcov"]
        w = 1.0 / (1.0 + np.exp(hyp[-1]))
        S_ = np.abs(theta[:, None, :] - theta[None, :, :]) / np.exp(hyp[:-1])
        R = w * np.prod(1 + S_, axis=2) * np.exp(-S_.sum(axis=2)) + (1 - w)
        Rn = (1 - nug) * R + nug * np.eye(n)
        ev, V = np.linalg.eigh(Rn)
        Vh = V / np.sqrt(np.abs(ev))
        pw = Vh @ (Vh.T @ G[:, k])
        emulist.append(dict(hypcov=hyp, hypind=hypind, nug=nug, Vh=Vh, pw=pw, sig2=float(G[:, k] @ pw / n)))
    resid = Z - G @ pct.T
    return dict(theta=theta, pct=pct, pcti=pct, scale=scale, offset=offset,
                extravar=np.mean(resid ** 2, axis=0) * scale ** 2, emulist=emulist, method="PCGP")
