"""Load the committed golden .npz files (tests/golden/*.npz, produced by make_golden.py from the
unmodified reference) and turn them into oracle state dicts."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL_CASES = ["c1_rbf", "c1_matern", "c1_logexp", "c1_nopca", "c1_multi", "odd_shape", "p20_trafo"]


def available():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
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


def oracle_states(g):
    """List of oracle state dicts (oracle/gp_oracle.py layout), one per emulator in golden `g`."""
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


def synthetic_module():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(
        "gpbt_synthetic", os.path.join(root, "gpbayestools-hic_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cov_exp_sys(g):
    """expdata_cov + the PSD systematic term make_golden.py added for lp_posterior_sys."""
    return g["cov_exp"] + synthetic_module().systematic_cov(g["cov_exp"].shape[0])
