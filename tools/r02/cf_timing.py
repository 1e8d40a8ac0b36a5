"""phase timing of the fused Cholesky from its clock64 stamps (option cf_debug): python tools/r02/cf_timing.py [N]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench, gpbt_b200
from gpbt_b200 import _lib, fixtures
from gpbt_b200.device import DeviceChain
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
g = fixtures.load("c2_rbf")
states, _ = fixtures.emulator_states(g)
ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"], devices=[0])
ch._checked = True
_lib.set_option("chol", "fused"); _lib.set_option("chol_batch", 1000000); _lib.set_option("chol_streams", 1)
X = torch.from_numpy(bench.walkers(g, N, 1)).cuda()
for _ in range(3):
    ch.log_target_device(X, -np.inf, path="dense")
torch.cuda.synchronize()
_lib.set_option("cf_debug", 1)
ch.log_target_device(X, -np.inf, path="dense")
torch.cuda.synchronize()
buf = np.zeros((16, 32, 8, 8), dtype=np.int64)
_lib.check(_lib.lib.gpbt_debug_timing_read(buf.ctypes.data, buf.nbytes))
_lib.set_option("cf_debug", None)
clk = 1.965e3  # cycles per us
for L in range(11):
    b = buf[L]
    t0 = b[:, 0]   # stamps: 0 start, 1 operand stream done, 2 Dinv landed (after the wait), 3 end
    ok = t0[:, 3] > 0
    if not ok.any():
        continue
    d = lambda a, c: np.median((t0[ok, c] - t0[ok, a]) / clk)
    line = "J=%4d tile0: ring %.1f wait %.1f solve+D %.1f" % (32 * (L - 1), d(0, 1), d(1, 2), d(2, 3))
    reg = b[:, 1:7, :]
    okr = reg[:, :, 3] > 0
    if okr.any():
        line += " | regular: ring %.1f wait %.1f solve %.1f (n=%d)" % (
            np.median((reg[..., 1] - reg[..., 0])[okr] / clk), np.median((reg[..., 2] - reg[..., 1])[okr] / clk),
            np.median((reg[..., 3] - reg[..., 2])[okr] / clk), okr.sum())
    f = b[:, 7]
    okf = f[:, 5] > 0
    if okf.any():
        df = lambda a, c: np.median((f[okf, c] - f[okf, a]) / clk)
        line += " | factor(Jd=%d): wait %.1f load %.1f chol %.1f inverse %.1f tail %.1f us" % (32 * L, df(0, 1), df(1, 2), df(2, 3), df(3, 4), df(4, 5))
    print(line)
