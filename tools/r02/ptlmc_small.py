"""PTLMC iteration at the reference's default ladder (50 temperatures + 16 chains) on the config-2 chain: host
loop (NumPy draws, one GPU call, C exchange helper) against the device-resident loop.
python tools/r02/ptlmc_small.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench, gpbt_b200
from gpbt_b200 import fixtures, ptlmc
from gpbt_b200.device import DeviceChain
g = fixtures.load("c2_rbf")
states, _ = fixtures.emulator_states(g)
ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"], devices=[0])
for n_hot, n_cold in ((50, 16), (500, 160)):
    n = n_hot + n_cold
    temps = ptlmc.temperature_ladder(n_hot, n_cold, 100.0)
    lo, hi = g["lo"], g["hi"]
    rng = np.random.default_rng(1)
    theta0 = 0.5 * (lo + hi) + 0.05 * (hi - lo) * rng.standard_normal((n, len(lo)))
    root = np.diag(0.02 * (hi - lo))
    theta = theta0.copy()
    f = ch.log_target(theta, -np.inf).reshape(-1, 1) / temps
    iters = 200
    np.random.seed(0)
    t0 = time.perf_counter()
    for _ in range(iters):
        prop = theta + np.sqrt(2) * temps ** (1 / 3) * (np.random.normal(0, 1, theta.shape) @ root)
        fp = ch.log_target(prop, -np.inf).reshape(-1, 1) / temps
        with np.errstate(invalid="ignore"):
            take = np.where(np.log(np.random.uniform(size=n)) < np.squeeze(fp - f))[0]
        theta[take], f[take] = prop[take], fp[take]
        flat = f * temps
        order = ptlmc.temp_exchange(flat, temps, iters=5)
        f, theta = flat[order] / temps, theta[order]
    us_host = (time.perf_counter() - t0) / iters * 1e6
    pt = ptlmc.DevicePTLMC(ch, temps, root, n_hot, seed=1)
    pt.set_state(theta0)
    pt.run(2 * iters, iters, n_steps=iters)
    t0 = time.perf_counter()
    pt.run(2 * iters, iters, n_steps=2 * iters)
    us_dev = (time.perf_counter() - t0) / (2 * iters) * 1e6
    out = pt.read()
    pt.close()
    print("config 2, %d + %d chains: host loop %.1f us per iteration, device loop %.1f us  (tau %.3f, %d accepted)" % (
        n_hot, n_cold, us_host, us_dev, out["tau"], out["accepted"]), flush=True)
ch.release()
