"""one dense-path call for ncu: python tools/r02/profile_dense.py N [chol option] [chol_batch]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench, gpbt_b200
from gpbt_b200 import _lib, fixtures
from gpbt_b200.device import DeviceChain
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = fixtures.load("c2_rbf")
states, _ = fixtures.emulator_states(g)
ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"], devices=[0])
ch._checked = True
if len(sys.argv) > 2:
    _lib.set_option("chol", sys.argv[2])
if len(sys.argv) > 3:
    _lib.set_option("chol_batch", sys.argv[3])
X = torch.from_numpy(bench.walkers(g, N, 1)).cuda()
for _ in range(int(os.environ.get("REPS", "2"))):
    lp = ch.log_target_device(X, -np.inf, path="dense")
torch.cuda.synchronize()
print("ok", float(lp[torch.isfinite(lp)].mean()))
