"""Short driver for ncu: kernel (a) kind 2 (surmise-PCGP-shaped emulator, config 3 shape) on 4096 chains.
    ncu --set full --import-source on --clock-control none -k regex:pc_predict -s 1 -c 1 -o out python tools/profile_band.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gpbt_b200  # noqa: E402,F401
from gpbt_b200 import synthetic  # noqa: E402
from gpbt_b200.device import DeviceEmulator  # noqa: E402
from gpbt_b200.state import EmulatorState  # noqa: E402

info = synthetic.pcgp_fitinfo(15, 1000, 300, 20)
de = DeviceEmulator(EmulatorState.from_pcgp_fitinfo(info))
lo, hi = synthetic.box(15)
X = torch.from_numpy(np.random.default_rng(0).uniform(lo, hi, (4096, 15))).cuda()
for _ in range(3):
    zm, zv = de.pc_predict_device(X)
torch.cuda.synchronize()
print("ok", float(zv.mean()))
