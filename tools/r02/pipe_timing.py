"""dense path at config 2 with kernel (a) pipelined over sub-batches (option chol_pipe):
python tools/r02/pipe_timing.py [N]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench, gpbt_b200
from gpbt_b200 import _lib, fixtures
from gpbt_b200.device import DeviceChain
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = fixtures.load("c2_rbf")
states, _ = fixtures.emulator_states(g)
X = torch.from_numpy(bench.walkers(g, N, 1)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(label, **opts):
    for k, v in opts.items():
        _lib.set_option(k, v)
    ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"], devices=[0])
    ch._checked = True
    for _ in range(3):
        lp = ch.log_target_device(X, -np.inf, path="dense")
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lp = ch.log_target_device(X, -np.inf, path="dense")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out = lp.cpu().numpy().copy()
    ch.release()
    for k in opts:
        _lib.set_option(k, None)
    print("%-28s median %.4f ms  min %.4f  -> %.3e evals/s" % (label, np.median(ts), np.min(ts), N / np.median(ts) * 1e3), flush=True)
    return out


ref = run("pipe off (2 streams)", chol_pipe=1)
run("default", )
for pipe in (2, 3, 4):
    for prio in (0,):
        for streams in (0, 3):
            got = run("pipe %d prio %d streams %d" % (pipe, prio, streams), chol_pipe=pipe, chol_prio=prio, chol_streams=streams)
            if not np.array_equal(got, ref):
                fin = np.isfinite(ref)
                print("   DIFFERS from pipe off: max |diff| %.3e" % np.max(np.abs(got[fin] - ref[fin])))
