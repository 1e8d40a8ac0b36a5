"""Run-to-run spread of the dense path (config 2, N = 2048): 40 calls, CUDA events per call."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import gpbt_b200  # noqa
from gpbt_b200.device import DeviceChain
from gpbt_b200.state import EmulatorState
g, sts = bench.load_c2()
states = [EmulatorState.from_arrays(s["kind"], s["Xtr"], s["ell"], s["c"], s["sn"], s["alpha"], s["mu"], s["scale"], s.get("A"), s.get("Ctrunc"), L=s["L"], keep_L=False) for s in sts]
chain = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
X = torch.from_numpy(bench.walkers(g, 2048, 1)).cuda()
for _ in range(3):
    chain.log_target_device(X, -np.inf, path="dense")
torch.cuda.synchronize()
ts = []
for _ in range(40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); chain.log_target_device(X, -np.inf, path="dense"); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = np.sort(ts)
print(os.environ.get("GPBT_CHOL", "default"), "min %.2f median %.2f p90 %.2f max %.2f ms" % (ts[0], np.median(ts), ts[35], ts[-1]))
