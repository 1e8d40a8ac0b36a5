"""Committed fixtures: the golden .npz files under tests/golden/ (written by tests/golden/make_golden.py
from the UNMODIFIED reference: trained hyper-parameters, inputs, outputs) as plain arrays, and the
emulator states built from them.  Shared by the tests, bench.py and the tools; no test code is imported."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def available():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    path = name if os.path.sep in name or name.endswith(".npz") else os.path.join(GOLDEN_DIR, name + ".npz")
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def rebuild_L(kind, Xtr, c, ell, sn, gpr_alpha=0.1):
    """L_ = cholesky(kernel_(Xtr) + alpha I) exactly as sklearn's fit does (_gpr.py:349-360);
    used when a golden file is too large to carry L_ (config 2)."""
    from scipy.linalg import cholesky
    from scipy.spatial.distance import pdist, squareform
    Xs = Xtr / ell
    if kind == "RBF":
        K = squareform(np.exp(-0.5 * pdist(Xs, metric="sqeuclidean")))
        np.fill_diagonal(K, 1.0)
    else:
        r = squareform(pdist(Xs, metric="euclidean")) * np.sqrt(3.0)
        K = (1.0 + r) * np.exp(-r)
    K = c * K
    K[np.diag_indices_from(K)] += sn       # WhiteKernel on the training diagonal
    K[np.diag_indices_from(K)] += gpr_alpha
    return cholesky(K, lower=True, check_finite=False)


def state_dicts(g):
    """One dict of plain arrays per emulator of golden `g` (kind, Xtr, ell, c, sn, alpha, L, mu, scale, A,
    Ctrunc, flags, optional pre-transform) -- the layout oracle/gp_oracle.py works on."""
    states = []
    for e in range(int(g["n_emu"])):
        pre = "e%d_" % e
        Xtr = g[pre + "Xtr"]
        n = Xtr.shape[0]
        q = g[pre + "alpha"].shape[0]
        kind = str(g[pre + "kind"])
        L = np.zeros((q, n, n))
        if pre + "Lpacked" in g:
            il = np.tril_indices(n)
            for j in range(q):
                L[j][il] = g[pre + "Lpacked"][j]
        else:
            for j in range(q):
                L[j] = rebuild_L(kind, Xtr, g[pre + "c"][j], g[pre + "ell"][j], g[pre + "sn"][j])
        st = dict(kind=kind, Xtr=Xtr, ell=g[pre + "ell"], c=g[pre + "c"], sn=g[pre + "sn"],
                  alpha=g[pre + "alpha"], L=L, no_pca=bool(g[pre + "no_pca"]),
                  exp_diag=bool(g[pre + "exp_diag"]), mu=g[pre + "mu"], scale=g[pre + "scale"])
        if not st["no_pca"]:
            st["A"] = g[pre + "A"]
            st["Ctrunc"] = g[pre + "Ctrunc"]
        if pre + "trafo_p_in" in g:
            grids = {"bulk": (0, (0.0, 0.5, 100)), "shear": (1, (0.0, 0.6, 100)), "yloss": (2, (0.0, 6.2, 100))}
            st["trafo"] = dict(p_in=int(g[pre + "trafo_p_in"]), groups=[
                dict(kind=grids[t][0], grid=grids[t][1], idx=g[pre + "trafo_%s_idx" % t],
                     smean=g[pre + "trafo_%s_smean" % t], sscale=g[pre + "trafo_%s_sscale" % t],
                     pmean=g[pre + "trafo_%s_pmean" % t], comp=g[pre + "trafo_%s_comp" % t])
                for t in ("bulk", "shear", "yloss")])
        states.append(st)
    return states




def emulator_states(g, keep_L=False):
    """(EmulatorState list for the device path, the plain-array dicts they were built from)"""
    from .state import EmulatorState, ParamTrafo
    dicts = state_dicts(g)
    out = []
    for s in dicts:
        trafo = None
        if "trafo" in s:
            trafo = ParamTrafo(p_in=s["trafo"]["p_in"], groups=s["trafo"]["groups"])
        out.append(EmulatorState.from_arrays(s["kind"], s["Xtr"], s["ell"], s["c"], s["sn"], s["alpha"], s["mu"],
                                             s["scale"], s.get("A"), s.get("Ctrunc"), L=s["L"], no_pca=s["no_pca"],
                                             exp_diag=s["exp_diag"], keep_L=keep_L, trafo=trafo))
    return out, dicts
