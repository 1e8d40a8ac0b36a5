"""The early-refill race of the fused Cholesky's operand ring: does a generic->async proxy fence in front of the
consumers' "slot is free" arrival cure it?  cf_debug = 16 * flags: 64 = refill after stage s - 1 (early),
128 = fence.proxy.async before the arrive.   python tools/r02/refill_experiment.py [reps]"""
import os, sys, time
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
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
for flags in (0, 64, 64 + 128, 128, 64, 64 + 128):
    for streams, cb in ((2, 2048), (2, 512), (3, 1024)):
        _lib.set_option("cf_debug", 16 * flags); _lib.set_option("chol_batch", cb); _lib.set_option("chol_streams", streams)
        bad = 0
        torch.cuda.synchronize(); t0 = time.time()
        for rep in range(reps):
            lp = ch.log_target_device(X, -np.inf, path="dense")
            d = (lp - ref).abs()
            d[~fin] = 0
            bad += int((d > 1e-8).sum())
        torch.cuda.synchronize()
        print("flags %3d (early %d fence %d) streams %d batch %4d: bad walkers %d of %.1fM  (%.2f s)" % (
            flags, bool(flags & 64), bool(flags & 128), streams, cb, bad, reps * 8192 / 1e6, time.time() - t0), flush=True)
