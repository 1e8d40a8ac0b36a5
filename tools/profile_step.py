"""Short driver for ncu: a few steps of the hot path on config 2 (lowrank path, N=4096) and one
dense-path step (N=512).  Usage: python tools/profile_step.py [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import gpbt_b200  # noqa: E402,F401
from gpbt_b200.device import DeviceChain  # noqa: E402
from gpbt_b200.state import EmulatorState  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dense_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 512
g, sts = bench.load_c2()
states = [EmulatorState.from_arrays(s["kind"], s["Xtr"], s["ell"], s["c"], s["sn"], s["alpha"], s["mu"],
                                    s["scale"], s.get("A"), s.get("Ctrunc"), L=s["L"], keep_L=False) for s in sts]
chain = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
X = torch.from_numpy(bench.walkers(g, 4096, 1)).cuda()
for _ in range(steps):
    lp = chain.log_target_device(X, -np.inf, path="lowrank")
if dense_rows:
    lpd = chain.log_target_device(X[:dense_rows], -np.inf, path="dense")
torch.cuda.synchronize()
print("ok", float(lp[torch.isfinite(lp)].mean()))
