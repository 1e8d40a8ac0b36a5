"""kernel (a) alone at config 2, N = 4096, for the walker-tile widths: python tools/r02/time_a.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench, gpbt_b200
from gpbt_b200 import _lib, fixtures
from gpbt_b200.device import DeviceEmulator
g = fixtures.load("c2_rbf")
states, _ = fixtures.emulator_states(g)
de = DeviceEmulator(states[0])
for N in (4096, 4736, 2368):
    Xs = [torch.from_numpy(bench.walkers(g, N, i)).cuda() for i in range(4)]
    for tile in (None, 8, 32, None):
        _lib.set_option("pc_tile", tile)
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
        fl = bench.flops_pc_predict(17, 500, 20) * N
        print("N %d pc_tile=%s: min %.4f ms median %.4f ms  -> %.2f TFLOP/s  (%.4f us per walker)" % (
            N, tile, min(ts), np.median(ts), fl / (min(ts) * 1e-3) / 1e12, min(ts) * 1e3 / N), flush=True)
_lib.set_option("pc_tile", None)
