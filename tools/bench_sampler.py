"""Sampler steps per second: the device-resident ensemble sampler (one CUDA graph per step) against a
host-side stretch-move loop that calls the GPU log-posterior once per half-ensemble, which is how
emcee drives the reference (pool=self).  Config 1 (128 walkers on the p5/n100/m50/q10 emulator) is
latency bound; config 2 (8192 walkers on p17/n500/m300/q20) is throughput bound.

    python tools/bench_sampler.py [out.jsonl]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpbt_b200  # noqa: E402,F401
from gpbt_b200 import _lib  # noqa: E402
from gpbt_b200.device import DeviceChain  # noqa: E402
from gpbt_b200.sampler import DeviceEnsembleSampler  # noqa: E402
from gpbt_b200.state import EmulatorState  # noqa: E402
from tests import goldens  # noqa: E402

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None


def emit(d):
    line = json.dumps(d)
    print(line)
    if out:
        out.write(line + "\n")
        out.flush()


def host_loop(dc, x0, steps, a=2.0, seed=0):
    """stretch move on the host, two GPU calls per step (the emcee pattern)"""
    rng = np.random.default_rng(seed)
    x = x0.copy()
    nw, p = x.shape
    half = nw // 2
    lp = dc.log_target(x, -np.inf)
    t0 = time.perf_counter()
    for _ in range(steps):
        for s, c in ((slice(0, half), slice(half, nw)), (slice(half, nw), slice(0, half))):
            ns = x[s].shape[0]
            z = ((a - 1.0) * rng.random(ns) + 1.0) ** 2 / a
            partner = x[c][rng.integers(0, x[c].shape[0], ns)]
            prop = partner + z[:, None] * (x[s] - partner)
            lp_new = dc.log_target(prop, -np.inf)
            acc = np.log(rng.random(ns)) < (p - 1) * np.log(z) + lp_new - lp[s]
            xs, ls = x[s].copy(), lp[s].copy()
            xs[acc], ls[acc] = prop[acc], lp_new[acc]
            x[s], lp[s] = xs, ls
    return steps / (time.perf_counter() - t0)


def device_rate(dc, x0, steps, use_graph):
    s = DeviceEnsembleSampler(x0.shape[0], x0.shape[1], dc, seed=1, use_graph=use_graph)
    s.set_state(x0)
    s.advance(min(steps, 50))                      # warm-up (workspaces, graph capture)
    l0 = _lib.lib.gpbt_launch_count()
    t0 = time.perf_counter()
    s.advance(steps)
    dt = time.perf_counter() - t0
    launches = (_lib.lib.gpbt_launch_count() - l0)
    af = float(s.acceptance_fraction.mean())
    s.close()
    return steps / dt, af, launches


def start(lo, hi, nw, seed=5):
    rng = np.random.default_rng(seed)
    mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo)
    return mid + 0.5 * half * rng.uniform(-1, 1, (nw, len(mid)))


def run(name, dc, lo, hi, nw, steps, host_steps):
    x0 = start(lo, hi, nw)
    for use_graph in (True, False):
        rate, af, launches = device_rate(dc, x0, steps, use_graph)
        emit({"config": name, "walkers": nw, "sampler": "device, " + ("CUDA graph per step" if use_graph else "eager launches"),
              "steps_per_s": rate, "evals_per_s": rate * nw, "us_per_step": 1e6 / rate, "acceptance": af,
              "kernel_launches_enqueued": int(launches), "steps": steps})
    rate = host_loop(dc, x0, host_steps)
    emit({"config": name, "walkers": nw, "sampler": "host stretch-move loop, 2 GPU calls per step (emcee pattern)",
          "steps_per_s": rate, "evals_per_s": rate * nw, "us_per_step": 1e6 / rate, "steps": host_steps})


g = goldens.load("c1_rbf")
sts = goldens.oracle_states(g)
states = [EmulatorState.from_arrays(s["kind"], s["Xtr"], s["ell"], s["c"], s["sn"], s["alpha"], s["mu"], s["scale"],
                                    s.get("A"), s.get("Ctrunc"), L=s["L"], keep_L=False) for s in sts]
dc = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
run("C1 (p5,n100,m50,q10)", dc, g["lo"], g["hi"], 128, 5000, 2000)
dc.release()

g2 = goldens.load("c2_rbf")
sts2 = goldens.oracle_states(g2)
states2 = [EmulatorState.from_arrays(s["kind"], s["Xtr"], s["ell"], s["c"], s["sn"], s["alpha"], s["mu"], s["scale"],
                                     s.get("A"), s.get("Ctrunc"), L=s["L"], keep_L=False) for s in sts2]
dc = DeviceChain(states2, g2["lo"], g2["hi"], g2["y_exp"].reshape(-1), g2["cov_exp"])
run("C2 (p17,n500,m300,q20)", dc, g2["lo"], g2["hi"], 8192, 300, 100)
dc.release()

# ---- config 3: one PTLMC iteration at 8192 chains on the surmise-PCGP-shaped emulator -------------
# (the move of gpbt_b200.ptlmc.sampler_ptlmc without its start-up: proposal, one GPU call, Metropolis,
# five exchange sweeps; the pre-optimiser is N = 1 L-BFGS-B calls per chain and is not timed here)
from gpbt_b200 import ptlmc, synthetic  # noqa: E402

info = synthetic.pcgp_fitinfo(15, 1000, 300, 20)
lo3, hi3 = synthetic.box(15)
y3 = info["offset"] + 0.3 * info["scale"]
dc = DeviceChain([EmulatorState.from_pcgp_fitinfo(info)], lo3, hi3, y3, np.diag((0.03 * np.abs(y3)) ** 2))
n_all = 8192
temps = ptlmc.temperature_ladder(n_all - 1024, 1024, 100.0)
theta = start(lo3, hi3, n_all)
f = dc.log_target(theta, -np.inf).reshape(-1, 1) / temps
root = np.diag(0.02 * (hi3 - lo3))
t_call = t_np = t_ex = t_ex_py = 0.0
iters = 10
np.random.seed(0)
for k in range(iters):
    t0 = time.perf_counter()
    prop = theta + np.sqrt(2) * temps ** (1 / 3) * (np.random.normal(0, 1, theta.shape) @ root)
    t1 = time.perf_counter()
    fp = dc.log_target(prop, -np.inf).reshape(-1, 1) / temps
    t2 = time.perf_counter()
    take = np.where(np.log(np.random.uniform(size=n_all)) < np.squeeze(fp - f))[0]
    theta[take], f[take] = prop[take], fp[take]
    t3 = time.perf_counter()
    flat = f * temps
    order = ptlmc.temp_exchange(flat, temps, iters=5)
    t4 = time.perf_counter()
    if k < 2:
        ptlmc.temp_exchange_python(flat, temps, iters=5)
        t_ex_py += time.perf_counter() - t4
    f, theta = flat[order] / temps, theta[order]
    t_np += (t1 - t0) + (t3 - t2)
    t_call += t2 - t1
    t_ex += t4 - t3
emit({"config": "C3 (p15,n1000,m300,q20, surmise-PCGP-shaped), PTLMC iteration at 8192 chains", "chains": n_all,
      "ms_gpu_call": 1e3 * t_call / iters, "ms_numpy_proposal_accept": 1e3 * t_np / iters,
      "ms_exchange_host_helper": 1e3 * t_ex / iters, "ms_exchange_python_loop": 1e3 * t_ex_py / 2,
      "iterations_per_s": iters / (t_call + t_np + t_ex)})
dc.release()
