"""How much accuracy the expanded squared distance |x|^2 + |t|^2 - 2 x.t would cost kernel (a) on the golden
cases (NumPy emulation; the kernel's own form is sum (x - t)^2): python tools/r02/dist_expansion_check.py"""
import os, sys
import numpy as np
from scipy.linalg import solve_triangular
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import goldens
from tests.helpers import var_tol, golden_tol

for case in [c for c in goldens.SMALL_CASES + ["c2_rbf", "c2_matern"] if c in goldens.available()]:
    g = goldens.load(case)
    Xin = g["X"][g["inside"]][:256]
    for e, st in enumerate(goldens.oracle_states(g)):
        if st["kind"] not in ("RBF", "Matern") or st.get("trafo") is not None:
            continue
        worst = 0.0, 0.0, 0.0
        for j in range(st["alpha"].shape[0]):
            xs, ts = Xin / st["ell"][j], st["Xtr"] / st["ell"][j]
            d_direct = ((xs[:, None, :] - ts[None, :, :]) ** 2).sum(-1)
            d_exp = np.maximum((xs * xs).sum(1)[:, None] + (ts * ts).sum(1)[None, :] - 2.0 * (xs @ ts.T), 0.0)
            out = []
            for d in (d_direct, d_exp):
                if st["kind"] == "RBF":
                    K = st["c"][j] * np.exp(-0.5 * d)
                else:
                    r = np.sqrt(d) * np.sqrt(3.0)
                    K = st["c"][j] * ((1.0 + r) * np.exp(-r))
                V = solve_triangular(st["L"][j], K.T, lower=True, check_finite=False)
                out.append((K @ st["alpha"][j], (st["c"][j] + st["sn"][j]) - np.einsum("ij,ij->j", V, V), K))
            (m0, v0, K0), (m1, v1, K1) = out
            tol = var_tol(v0[:, None], st["c"][j:j + 1], st["sn"][j:j + 1])[:, 0]
            scale = (np.abs(K0) @ np.abs(st["alpha"][j])).max()
            worst = (max(worst[0], np.max(np.abs(v1 - v0) / tol)), max(worst[1], np.max(np.abs(m1 - m0)) / (1e-12 * max(scale, 1.0) * 10)),
                     max(worst[2], np.max(np.abs(K1 - K0) / np.maximum(np.abs(K0), 1e-300))))
        print("%-12s emu %d  scaled |x|^2 up to %.1f:  var err / var_tol %.3g   mean err / mean_tol %.3g   K rel err %.2e" % (
            case, e, (xs * xs).sum(1).max(), worst[0], worst[1], worst[2]))
