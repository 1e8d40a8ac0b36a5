"""values of the fused dense path at large N / several streams against the low-rank path"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench, gpbt_b200
from gpbt_b200 import _lib, fixtures
from gpbt_b200.device import DeviceChain
g = fixtures.load("c2_rbf")
states, _ = fixtures.emulator_states(g)
ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"], devices=[0])
X = torch.from_numpy(bench.walkers(g, 8192, 1)).cuda()
ref = ch.log_target_device(X, -np.inf, path="lowrank")
fin = torch.isfinite(ref)
_lib.set_option("chol", "fused")
def count(label, reps, N=8192):
    tot_bad = 0
    torch.cuda.synchronize()
    t0 = time.time()
    for rep in range(reps):
        lp = ch.log_target_device(X[:N], -np.inf, path="dense")
        d = (lp - ref[:N]).abs()
        d[~fin[:N]] = 0
        tot_bad += int((d > 1e-8).sum())
    torch.cuda.synchronize()
    print("%-44s bad walkers %d of %.1fM  (%.2f s)" % (label, tot_bad, reps * N / 1e6, time.time() - t0), flush=True)

_lib.set_option("chol", "fused")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
import time
for pipe in (2, 4):
    _lib.set_option("chol_pipe", pipe)
    count("fused pipelined over %d sub-batches" % pipe, reps)
_lib.set_option("chol_pipe", 1)
for dbg in (0, 16):
    for streams, cb in ((1, 100000), (1, 512), (2, 2048), (2, 512), (3, 1024)):
        _lib.set_option("cf_debug", dbg); _lib.set_option("chol_batch", cb); _lib.set_option("chol_streams", streams)
        count("fused flags %d streams %d batch %d" % (dbg >> 4, streams, cb), reps)
