"""fused (b)+(c) Cholesky against the low-rank path and the materialised dense path; timing sweep"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench, gpbt_b200
from gpbt_b200 import _lib, fixtures
from gpbt_b200.device import DeviceChain

def chain_of(name, cov=None):
    g = fixtures.load(name)
    states, _ = fixtures.emulator_states(g)
    return g, DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"] if cov is None else cov, devices=[0])

for name in ("c1_rbf", "odd_shape", "c1_multi", "c2_rbf"):
    g, ch = chain_of(name)
    rng = np.random.default_rng(1)
    X = rng.uniform(g["lo"], g["hi"], (700, len(g["lo"])))
    X[::41, 0] = g["hi"][0] + 1
    ref = ch.log_target(X, -np.inf, path="lowrank")
    _lib.set_option("chol", "batch")
    old = ch.log_target(X, -np.inf, path="dense")
    _lib.set_option("chol", "fused")
    new = ch.log_target(X, -np.inf, path="dense")
    _lib.set_option("chol", None)
    fin = np.isfinite(ref)
    print(name, "M", ch.M, "fused vs lowrank %.3e  stepped vs lowrank %.3e  inf pattern %s notpd %d" % (
        np.max(np.abs(new[fin] - ref[fin])), np.max(np.abs(old[fin] - ref[fin])),
        np.array_equal(np.isfinite(new), fin), ch.last_notpd), flush=True)
    ch.release()

g, ch = chain_of("c2_rbf")
X = torch.from_numpy(bench.walkers(g, 8192, 1)).cuda()
def t(N, reps=5):
    ch.log_target_device(X[:N], -np.inf, path="dense"); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ch.log_target_device(X[:N], -np.inf, path="dense"); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
def best(N, reps=7):
    for _ in range(2):
        ch.log_target_device(X[:N], -np.inf, path="dense")
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ch.log_target_device(X[:N], -np.inf, path="dense"); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.min(ts)), float(np.median(ts))
_lib.set_option("chol", "batch")
print("stepped       ", " ".join("N=%d: min %.3f med %.3f ms (%.2f us/w)" % ((N,) + best(N) + (best(N)[0] * 1e3 / N,)) for N in (1024, 4096, 8192)), flush=True)
_lib.set_option("chol", "fused")
for streams in (1, 2, 3, 4):
    for cb in (512, 1024, 2048, 100000):
        _lib.set_option("chol_batch", cb); _lib.set_option("chol_streams", streams)
        print("fused streams %d batch %6d" % (streams, cb), " ".join("N=%d: min %.3f med %.3f ms (%.2f us/w)" % ((N,) + best(N) + (best(N)[0] * 1e3 / N,)) for N in (1024, 4096, 8192)), flush=True)
_lib.set_option("chol", None); _lib.set_option("chol_batch", None); _lib.set_option("chol_streams", None)
