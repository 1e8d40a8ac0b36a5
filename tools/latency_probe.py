"""Where do the ~52 us of a small-N Chain.log_posterior call go?  (config 1, N = 64)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gpbt_b200  # noqa
from gpbt_b200 import _lib
from gpbt_b200.device import DeviceChain
from tests import goldens
from tests.helpers import product_states

g = goldens.load("c1_rbf")
states, _ = product_states(g)
ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
X = np.ascontiguousarray(g["X"][:64])
ch.log_target(X, -np.inf)

def timeit(f, n=2000):
    for _ in range(50): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6

print("host call (H2D + 2 kernels + D2H + sync)    %.1f us" % timeit(lambda: ch.log_target(X, -np.inf)))
Xd = torch.from_numpy(X).cuda(); lpd = torch.empty(64, dtype=torch.float64, device="cuda")
print("device call, no sync (launch cost only)     %.1f us" % timeit(lambda: ch.log_target_device(Xd, -np.inf, lp_d=lpd)))
def dev_sync():
    ch.log_target_device(Xd, -np.inf, lp_d=lpd); torch.cuda.synchronize()
print("device call + sync (kernel latency)         %.1f us" % timeit(dev_sync))
Xp = torch.from_numpy(X).pin_memory().numpy()
print("host call from pinned X                     %.1f us" % timeit(lambda: ch.log_target(Xp, -np.inf)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): ch.log_target_device(Xd, -np.inf, lp_d=lpd)
e1.record(); torch.cuda.synchronize()
print("GPU time per step (events, back to back)    %.1f us" % (e0.elapsed_time(e1) * 1e3 / 200))
