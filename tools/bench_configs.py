"""Throughput of every BASELINE.json configuration on one GPU (bench.py times config 2 only; the
others are parity-test shapes -- this tool records what they do).  Writes one JSON object per line.
Spot checks compare against the committed golden vectors of the reference (never against oracle/:
only tests/, smoke() and bench.py's CPU-baseline legs may use it).

    python tools/bench_configs.py [out.jsonl]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import gpbt_b200  # noqa: E402,F401
from gpbt_b200 import synthetic  # noqa: E402
from gpbt_b200.device import DeviceChain, DeviceEmulator  # noqa: E402
from gpbt_b200.state import EmulatorState  # noqa: E402
from tests import goldens  # noqa: E402

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None


def emit(d):
    line = json.dumps(d)
    print(line)
    if out:
        out.write(line + "\n")
        out.flush()


def host_rate(chain, X, reps, path=None):
    chain.log_target(X, -np.inf, path=path)
    torch.cuda.synchronize()
    times = []
    for _ in range(max(reps, 5)):
        t0 = time.perf_counter()
        lp = chain.log_target(X, -np.inf, path=path)
        times.append(time.perf_counter() - t0)
    dt = float(np.median(times))   # (shared boxes show occasional 10+ ms stalls; the median ignores them)
    return len(X) / dt, dt, lp


def states_of(g):
    sts = goldens.oracle_states(g)
    return [EmulatorState.from_arrays(s["kind"], s["Xtr"], s["ell"], s["c"], s["sn"], s["alpha"], s["mu"], s["scale"],
                                      s.get("A"), s.get("Ctrunc"), L=s["L"], keep_L=False) for s in sts], sts


# ---- config 1: p5 n100 m50 q10, emcee 128 walkers (half-ensemble calls of 64) ---------------------
g = goldens.load("c1_rbf")
states, sts = states_of(g)
ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
for N in (1, 64, 128):
    X = np.ascontiguousarray(g["X"][:N])
    rate, dt, lp = host_rate(ch, X, 200)
    fin = np.isfinite(g["lp_posterior"][:N])
    emit({"config": "C1 (p5,n100,m50,q10)", "N_per_call": N, "evals_per_s": rate, "us_per_call": dt * 1e6,
          "max_abs_diff_vs_reference": float(np.max(np.abs(lp[fin] - g["lp_posterior"][:N][fin]))) if fin.any() else None,
          "note": "host API (Chain.log_posterior), one call per emcee half-step; bounded by launch + PCIe latency"})
# a large ensemble on the small emulator (throughput of the n = 100 shape, device-timed)
Xb = torch.from_numpy(np.ascontiguousarray(np.tile(g["X"], (512, 1)))).cuda()
ch.log_target_device(Xb, -np.inf)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ch.log_target_device(Xb, -np.inf)
e1.record()
torch.cuda.synchronize()
emit({"config": "C1 (p5,n100,m50,q10)", "N_per_call": Xb.shape[0], "evals_per_s": Xb.shape[0] * 10 / (e0.elapsed_time(e1) * 1e-3),
      "note": "device resident, large ensemble on the small emulator"})
del Xb
ch.release()

# ---- config 2: p17 n500 m300 q20, pocoMC 4096 particles -----------------------------------------
g2 = goldens.load("c2_rbf")
states2, sts2 = states_of(g2)
ch2 = DeviceChain(states2, g2["lo"], g2["hi"], g2["y_exp"].reshape(-1), g2["cov_exp"])
X = bench.walkers(g2, 4096, 5)
for path in ("lowrank", "dense"):
    rate, dt, lp = host_rate(ch2, X, 20 if path == "lowrank" else 3, path)
    emit({"config": "C2 (p17,n500,m300,q20)", "N_per_call": 4096, "path": path, "evals_per_s": rate,
          "ms_per_call": dt * 1e3})
lpg = ch2.log_target(g2["X"], -np.inf)
fin = np.isfinite(g2["lp_posterior"])
emit({"config": "C2", "check": "gpu vs reference golden, %d rows" % fin.sum(),
      "max_abs_diff": float(np.max(np.abs(lpg[fin] - g2["lp_posterior"][fin])))})

# ---- config 4: C2 state, large batches, full (non-diagonal) experimental covariance --------------
cov_sys = g2["cov_exp"] + synthetic.systematic_cov(300)
ch4 = DeviceChain(states2, g2["lo"], g2["hi"], g2["y_exp"].reshape(-1), cov_sys)
for N in (1 << 17, 1 << 20):
    X = bench.walkers(g2, N, 6)
    rate, dt, lp = host_rate(ch4, X, 2)
    emit({"config": "C4 (C2 state, full Sigma_exp)", "N_per_call": N, "evals_per_s": rate, "ms_per_call": dt * 1e3,
          "finite_fraction": float(np.isfinite(lp).mean())})
lpg = ch4.log_target(g2["X"], -np.inf)
emit({"config": "C4", "check": "gpu vs reference golden (full Sigma_exp), %d rows" % fin.sum(),
      "max_abs_diff": float(np.max(np.abs(lpg[fin] - g2["lp_posterior_sys"][fin])))})
ch4.release()

# ---- config 5: posterior-predictive sweep, Emulator.predict over LHD points ----------------------
de = DeviceEmulator(states2[0])
Npts = 1 << 21
Xd = torch.from_numpy(bench.walkers(g2, 1 << 18, 8)).cuda()
de.predict_diag_device(Xd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(Npts // Xd.shape[0]):
    mean, var = de.predict_diag_device(Xd)
e1.record()
torch.cuda.synchronize()
emit({"config": "C5 (C2 state) predict mean + diag(cov)", "points": Npts, "points_per_s": Npts / (e0.elapsed_time(e1) * 1e-3),
      "note": "device resident, 2^18-row chunks; the 10M-point sweep is 10M / this rate per GPU"})
rows = 8192
for _ in range(2):   # warm the caching allocator (2 x 5.9 GB) before timing
    mean, cov = de.predict_device(Xd[:rows], True)
torch.cuda.synchronize()
e0.record()
for _ in range(4):
    mean, cov = de.predict_device(Xd[:rows], True)
e1.record()
torch.cuda.synchronize()
dt = e0.elapsed_time(e1) * 1e-3 / 4
emit({"config": "C5 (C2 state) predict(return_cov=True), covariance materialised in HBM", "rows_per_chunk": rows,
      "points_per_s": rows / dt, "cov_write_GBps": rows * 300 * 300 * 8 / dt / 1e9})
del mean, cov
ch2.release()

# ---- config 3 shape: p15 n1000 (m300, q20), 8192 chains -- sklearn-kernel emulator of that shape ----
arr = synthetic.untrained_state_arrays(15, 1000, 300, 20)
st3 = EmulatorState.from_arrays(**arr, keep_L=False)
lo, hi = synthetic.box(15)
y_exp = synthetic.Simulator(15, 300)(lo + 0.4 * (hi - lo))[0]
cov_exp = np.diag((0.03 * np.abs(y_exp)) ** 2)
ch3 = DeviceChain([st3], lo, hi, y_exp, cov_exp)
X = synthetic.walkers(15, 8192, seed=9)
rate, dt, lp = host_rate(ch3, X, 5)
dense = ch3.log_target(X[:256], -np.inf, path="dense")
fin = np.isfinite(dense)
emit({"config": "C3 shape (p15,n1000,m300,q20), RBF GP emulator (surmise PCSK itself: oracle absent, unpinned)",
      "N_per_call": 8192, "evals_per_s": rate, "ms_per_call": dt * 1e3,
      "max_abs_diff_lowrank_vs_dense_256_rows": float(np.max(np.abs(lp[:256][fin] - dense[fin]))),
      "note": "parity of this shape against the oracle: tests/test_gpu_parity.py::test_n1000_shape"})
ch3.release()

# ---- config 3 proper: surmise-PCGP-shaped emulator (kernel kind 2: separable Matern + constant, dense Vh) ----
info = synthetic.pcgp_fitinfo(15, 1000, 300, 20)
stb = EmulatorState.from_pcgp_fitinfo(info)
yb = info["offset"] + 0.3 * info["scale"]
chb = DeviceChain([stb], lo, hi, yb, np.diag((0.03 * np.abs(yb)) ** 2))
rate, dt, lp = host_rate(chb, X, 5)
dense = chb.log_target(X[:256], -np.inf, path="dense")
fin = np.isfinite(dense)
deb = DeviceEmulator(stb)
Xd3 = torch.from_numpy(X).cuda()
deb.pc_predict_device(Xd3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    deb.pc_predict_device(Xd3)
e1.record()
torch.cuda.synchronize()
ka = e0.elapsed_time(e1) * 1e-3 / 5
# per evaluation: q n (4 p + 6) for the kernel row (|d|, sum, product per dimension) + 2 q n for the mean
# + 2 q n^2 for r Vh (dense) + 2 q n for the squares
fl = 20 * 1000 * (4 * 15 + 6) + 2 * 20 * 1000 + 2 * 20 * 1000 * 1000 + 2 * 20 * 1000
emit({"config": "C3 (p15,n1000,m300,q20), surmise-PCGP-shaped emulator (EmulatorBAND path; parity unpinned: surmise absent)",
      "N_per_call": 8192, "evals_per_s": rate, "ms_per_call": dt * 1e3,
      "kernel_a_ms": ka * 1e3, "kernel_a_TFLOPs": fl * 8192 / ka / 1e12, "algorithmic_flops_per_eval": fl,
      "max_abs_diff_lowrank_vs_dense_256_rows": float(np.max(np.abs(lp[:256][fin] - dense[fin]))),
      "note": "values against the restatement: tests/test_gpu_band.py"})
chb.release()
