"""dense path at config 2, a few option sets, for A/B runs of library builds (GPBT_B200_LIB):
python tools/r02/ab_timing.py N "k=v,k=v" "k=v" ..."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench, gpbt_b200
from gpbt_b200 import _lib, fixtures
from gpbt_b200.device import DeviceChain
N = int(sys.argv[1])
g = fixtures.load("c2_rbf")
states, _ = fixtures.emulator_states(g)
X = torch.from_numpy(bench.walkers(g, N, 1)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for spec in sys.argv[2:] or [""]:
    opts = dict(kv.split("=") for kv in spec.split(",") if kv)
    for k, v in opts.items():
        _lib.set_option(k, v)
    ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"], devices=[0])
    ch._checked = True
    for _ in range(3):
        lp = ch.log_target_device(X, -np.inf, path="dense")
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lp = ch.log_target_device(X, -np.inf, path="dense")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out = lp.cpu().numpy().copy()
    if ref is None:
        ref = ch.log_target_device(X, -np.inf, path="lowrank").cpu().numpy()
    fin = np.isfinite(ref)
    ch.release()
    for k in opts:
        _lib.set_option(k, None)
    print("%-40s median %.4f ms  min %.4f  -> %.3e evals/s   max|d| vs lowrank %.2e" % (
        os.path.basename(os.environ.get("GPBT_B200_LIB", "lib")) + " " + (spec or "default"), np.median(ts), np.min(ts),
        N / np.median(ts) * 1e3, np.max(np.abs(out[fin] - ref[fin]))), flush=True)
