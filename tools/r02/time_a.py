"""kernel (a) alone at config 2, N = 4096: python tools/r02/time_a.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench, gpbt_b200
from gpbt_b200 import fixtures
from gpbt_b200.device import DeviceEmulator
g = fixtures.load("c2_rbf")
states, _ = fixtures.emulator_states(g)
de = DeviceEmulator(states[0])
Xs = [torch.from_numpy(bench.walkers(g, 4096, i)).cuda() for i in range(4)]
for i in range(5):
    de.pc_predict_device(Xs[i % 4])
torch.cuda.synchronize()
ts = []
for rep in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20):
        zm, zv = de.pc_predict_device(Xs[i % 4])
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) / 20)
fl = bench.flops_pc_predict(17, 500, 20) * 4096
print("kernel (a): min %.4f ms median %.4f ms  -> %.2f TFLOP/s" % (min(ts), np.median(ts), fl / (min(ts) * 1e-3) / 1e12))
print("checksum", float(zm.sum()), float(zv.sum()))
