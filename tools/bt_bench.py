"""kernel (b) alone: predict_device(return_cov=True) minus kernel (a), config-2 state, N = 1024 / 4096."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import gpbt_b200  # noqa
from gpbt_b200 import _lib
from gpbt_b200.device import DeviceEmulator, _stream_ptr
from gpbt_b200.state import EmulatorState
g, sts = bench.load_c2()
s = sts[0]
st = EmulatorState.from_arrays(s["kind"], s["Xtr"], s["ell"], s["c"], s["sn"], s["alpha"], s["mu"], s["scale"], s.get("A"), s.get("Ctrunc"), L=s["L"], keep_L=False)
de = DeviceEmulator(st)
for N in (1024, 4096):
    X = torch.from_numpy(bench.walkers(g, N, 1)).cuda()
    zm, zv = de.pc_predict_device(X)
    mean = torch.empty((N, st.m), dtype=torch.float64, device="cuda")
    cov = torch.empty((N, st.m, st.m), dtype=torch.float64, device="cuda")
    def run():
        _lib.check(_lib.lib.gpbt_backtransform(st.handle(), zm.data_ptr(), zv.data_ptr(), st.q, mean.data_ptr(), st.m, cov.data_ptr(), st.m, 0, N, _stream_ptr(torch)))
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print(N, "ms", round(t, 4), "write TB/s", round(N * 300 * 300 * 8 / t / 1e9, 2), "sym check", float((cov - cov.transpose(1, 2)).abs().max()))
