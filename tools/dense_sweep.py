import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, gpbt_b200
from gpbt_b200.device import DeviceChain
from gpbt_b200.state import EmulatorState
g, sts = bench.load_c2()
states = [EmulatorState.from_arrays(s["kind"], s["Xtr"], s["ell"], s["c"], s["sn"], s["alpha"], s["mu"], s["scale"], s.get("A"), s.get("Ctrunc"), L=s["L"], keep_L=False) for s in sts]
ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
X = torch.from_numpy(bench.walkers(g, 8192, 1)).cuda()
for N in (512, 1024, 2048, 4096, 2048, 1024, 8192):
    ch.log_target_device(X[:N], -np.inf, path="dense"); torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ch.log_target_device(X[:N], -np.inf, path="dense"); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(N, ["%.2f" % t for t in ts], "us/walker %.2f" % (min(ts) * 1e3 / N))
