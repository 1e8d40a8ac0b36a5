"""Small end-to-end pass over every kernel (all paths, odd shapes) for compute-sanitizer.
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpbt_b200
from gpbt_b200 import _lib  # noqa: E402,F401
from gpbt_b200.device import DeviceChain, mvn_loglike_batch  # noqa: E402
from gpbt_b200.emulator import Emulator  # noqa: E402
from tests import goldens  # noqa: E402
from tests.helpers import product_states  # noqa: E402

for case in ("odd_shape", "c1_multi", "c1_nopca", "c1_logexp", "p20_trafo", "c1_matern"):
    g = goldens.load(case)
    states, _ = product_states(g)
    ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
    X = g["X"]
    ref = g["lp_posterior"]
    fin = np.isfinite(ref)
    for path in ("auto", "dense"):
        for tile in ("8", "16", "32"):
            _lib.set_option("pc_tile", tile)
            lp = ch.log_target(X, -np.inf, path=path)
            assert np.max(np.abs(lp[fin] - ref[fin])) <= 1e-8, (case, path, tile)
    _lib.set_option("pc_tile", None)
    for which in ("warp", "cta", "staged"):
        _lib.set_option("chol", which)
        lp = ch.log_target(X[:9], -np.inf, path="dense")
        mean, cov = ch.predict(X[g["inside"]][:5], 0.05)
    _lib.set_option("chol", None)
    emu = Emulator.from_state(states[0])
    emu.predict(X[g["inside"]][:7], return_cov=True, extra_std=0.1)
    emu.predict_diag(X[g["inside"]][:7])
    ch.release()
    print(case, "ok")
print("done")
