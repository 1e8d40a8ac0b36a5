#!/usr/bin/env python
"""
bench.py -- log-posterior evaluations / second for BASELINE.json's config 2
(17 parameters, 500 design points, 300 observables, 20 PCs; 4096 particles per GPU per step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (Chain.log_posterior) over one fresh batch of walkers.
  value     device-timed (CUDA events on the launching stream), inputs resident in HBM,
            L2 flushed between timed steps (outside the events)
  e2e       the same metric through the host API with HOST buffers (pinned): H2D of X, kernels,
            (all-gather), D2H of lp, one synchronisation per step -- wall clock
  roofline  kernel (a) pc_predict alone: algorithmic FP64 flops / CUDA-event time vs. the cuBLAS
            DGEMM rate measured in this run (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline  the oracle port (NumPy/SciPy restatement of the reference) on a bounded sample,
            on the host's cores, rank 0, N=1 only
--impl reference times the CPU restatement of the reference on the host cores (the reference itself
is pure Python that cannot travel to the GPU box).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "log-posterior evals/sec (walkers x steps)"
UNIT = "evals/s"
WORKLOAD = "C2: 17-param JETSCAPE-like design, 500 pts, 300 observables, 20 PCs; pocoMC 4096 particles"
SHAPE = dict(p=17, n=500, m=300, q=20)


def flops_pc_predict(p, n, q):
    """q n (3p+4) + 2 q n + q n^2 FP64 flops per evaluation (FMA = 2)"""
    return q * n * (3 * p + 4) + 2 * q * n + q * n * n


def load_c2():
    """Config-2 emulator state: hyper-parameters / alpha / PCA matrices trained by the unmodified
    reference (tests/golden/c2_rbf.npz); L_ is rebuilt from them as sklearn's fit does."""
    from tests import goldens
    g = goldens.load("c2_rbf")
    return g, goldens.oracle_states(g)


def walkers(g, N, seed):
    """uniform in the box, 1 % of rows pushed out of bounds"""
    rng = np.random.default_rng(seed)
    lo, hi = g["lo"], g["hi"]
    X = rng.uniform(lo, hi, (N, len(lo)))
    rows = rng.choice(N, max(1, N // 100), replace=False)
    X[rows, rng.integers(0, len(lo), len(rows))] = hi.max() + 1.0
    return X


class ClockSampler:
    """SM clock and throttle reasons sampled (NVML, every ~5 ms) while the timed region runs"""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except (ValueError, IndexError):
                pass
        return index

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                break
            time.sleep(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz,
                "reasons": [k for k, bit in names.items() if self.reasons & bit], "samples": len(self.sm)}


# ---- CPU baseline: the oracle port on all host cores ---------------------------------------------
# A pool of single-BLAS-thread worker processes over 128-row chunks -- the reference's own
# recommended deployment (examples/RunBayesianAnalysis.ipynb: pool = 12 with OMP_NUM_THREADS=1).
_CPU = {}


def _cpu_init():
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        _CPU["limit"] = threadpool_limits(1)
    except Exception:
        pass
    _CPU["g"], _CPU["sts"] = load_c2()


def _cpu_chunk(X):
    from oracle import gp_oracle as orc
    g = _CPU["g"]
    return orc.log_posterior(_CPU["sts"], X, g["lo"], g["hi"], g["y_exp"], g["cov_exp"])


class CpuPort:
    def __init__(self, workers=None):
        import multiprocessing as mp
        self.workers = workers or os.cpu_count()
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_cpu_init)
        self.pool.map(_cpu_chunk, [walkers(load_c2()[0], 2, 7)] * self.workers)   # warm every worker

    def rate(self, g, rows, chunk=128, seed=99):
        X = walkers(g, rows, seed)
        parts = [X[s:s + chunk] for s in range(0, rows, chunk)]
        t0 = time.perf_counter()
        out = self.pool.map(_cpu_chunk, parts, chunksize=1)
        dt = time.perf_counter() - t0
        return rows / dt, dt, np.concatenate(out)

    def close(self):
        self.pool.terminate()
        self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g, sts = load_c2()
    rows = args.cpu_rows
    port = CpuPort()
    for _ in range(max(1, min(args.warmup, 2))):
        port.rate(g, min(rows, 8 * port.workers))
    total = 0.0
    for s in range(args.steps):
        total += port.rate(g, rows, seed=100 + s)[1]
    port.close()
    value = rows * args.steps / total
    sample = ("%d walkers per step in 128-row chunks over a pool of %d single-BLAS-thread processes "
              "(NumPy/SciPy oracle port, O(N) per row)" % (rows, port.workers))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_step": rows,
                   "note": "CPU restatement of the reference: the reference itself is pure Python outside the repo and "
                           "cannot travel to the GPU box; its own sklearn call is O(N^2) per batch and slower than this port"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": port.workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gpbt_b200  # noqa: F401
    from gpbt_b200 import _lib
    from gpbt_b200.device import DeviceChain, DeviceEmulator
    from gpbt_b200.dist import ShardedEvaluator
    from gpbt_b200.state import EmulatorState

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved_stdout = None
    if world > 1:
        # NCCL prints a version banner on stdout when its first communicator is created (whatever
        # NCCL_DEBUG says on this image): send everything but the final JSON line to stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    g, sts = load_c2()
    states = [EmulatorState.from_arrays(s["kind"], s["Xtr"], s["ell"], s["c"], s["sn"], s["alpha"], s["mu"],
                                        s["scale"], s.get("A"), s.get("Ctrunc"), L=s["L"], keep_L=False) for s in sts]
    chain = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
    N, p = args.walkers, len(g["lo"])
    K, W = args.steps, args.warmup
    nb = K + W
    # fresh walkers every step; each rank has its own rows (weak scaling)
    Xh = [torch.from_numpy(walkers(g, N, 1000 * rank + i)).pin_memory() for i in range(nb)]
    Xd = [x.to(dev) for x in Xh]
    ev = ShardedEvaluator(lambda X: chain.log_target_device(X, -np.inf, path=args.path), dev)
    # multi-GPU: the all-gather of lp is fused into the last kernel (NVLink peer stores + one device
    # barrier); NCCL all-gather is the fallback (--collective nccl, or symmetric memory unavailable)
    collective = "none"
    gather_fn = ev.evaluate_local
    if world > 1:
        collective = "nccl all-gather of lp (8 B/walker)"
        if args.collective == "fused":
            try:
                from gpbt_b200.dist import PeerGather
                pg = PeerGather(N, dev)
                gather_fn = lambda X: pg.evaluate(chain, X, -np.inf, path=args.path)   # noqa: E731
                collective = "fused: peer stores from the last kernel over NVLink + device barrier"
            except Exception as exc:   # symmetric memory not available on this box
                collective += " (fused unavailable: %s)" % type(exc).__name__
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg -------------------------------------------------------------
    for i in range(W):
        gather_fn(Xd[i])
    barrier()
    launches0 = _lib.lib.gpbt_launch_count()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    with ClockSampler(local) as clk:
        t_wall0 = time.perf_counter()
        for i in range(K):
            flush.fill_(i & 0xff)
            ev0[i].record()
            out = gather_fn(Xd[W + i])
            ev1[i].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = _lib.lib.gpbt_launch_count() - launches0
    dev_ms = maxr(sum(a.elapsed_time(b) for a, b in zip(ev0, ev1)))
    value = world * N * K / (dev_ms * 1e-3)
    lp_check = out[:N].cpu().numpy()

    # ---- end-to-end leg: host buffers in, host result out ---------------------------------
    lp_host = torch.empty(world * N, dtype=torch.float64).pin_memory()

    def e2e_step(i):
        if world == 1:
            return chain.log_target(Xh[i].numpy(), -np.inf, path=args.path)      # gpbt_log_posterior_host
        xd = Xh[i].to(dev, non_blocking=True)
        lp_host.copy_(gather_fn(xd), non_blocking=True)
        torch.cuda.synchronize()
        return lp_host.numpy()

    for i in range(W):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        e2e_out = e2e_step(W + i)
    barrier()
    e2e_s = maxr(time.perf_counter() - t0)
    e2e_value = world * N * K / e2e_s

    # ---- dominant kernel alone: (a) pc_predict ---------------------------------------------
    de = DeviceEmulator(states[0])
    for i in range(3):
        de.pc_predict_device(Xd[i % nb])
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(K, 10)
    a0.record()
    for i in range(reps):
        de.pc_predict_device(Xd[i % nb])
    a1.record()
    torch.cuda.synchronize()
    ka_ms = a0.elapsed_time(a1) / reps
    fl = flops_pc_predict(SHAPE["p"], SHAPE["n"], SHAPE["q"]) * N
    achieved = fl / (ka_ms * 1e-3) / 1e12

    line = None
    if rank == 0:
        # DRAM traffic of the dominant kernel per launch, from the committed ncu capture
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_pc_predict_traffic.json")) as fh:
                tj = json.load(fh)
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except (OSError, KeyError, ValueError):
            pass
        # HBM denominator: the driver's measured copy rate, else the profiling guide's fallback
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                hbm_peak, hbm_src = float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json"
        except (OSError, KeyError, ValueError):
            pass
        # dense path ((a) -> (b) cov in HBM -> (c)), reported separately; timed before the
        # DGEMM below, whose power draw lowers the clocks of whatever runs right after it
        dense = None
        if args.dense_steps > 0:
            Nd = args.dense_walkers
            for i in range(3):
                chain.log_target_device(Xd[i % nb][:Nd], -np.inf, path="dense")
            torch.cuda.synchronize()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            for i in range(args.dense_steps):
                lpd = chain.log_target_device(Xd[i % nb][:Nd], -np.inf, path="dense")
            d1.record()
            torch.cuda.synchronize()
            dms = d0.elapsed_time(d1) / args.dense_steps
            ref = chain.log_target_device(Xd[(args.dense_steps - 1) % nb][:Nd], -np.inf, path="lowrank")
            fin = torch.isfinite(ref)
            dense = {"value": Nd / (dms * 1e-3), "unit": UNIT, "walkers": Nd, "ms_per_step": dms,
                     "max_abs_diff_vs_lowrank": float((lpd[fin] - ref[fin]).abs().max().item())}
        # FP64 roofline denominator: cuBLAS DGEMM, measured here
        n = 8192 if args.dgemm else 0
        peak, peak_src = 35.45, "profiles/r01_dgemm_peak.json (cuBLAS DGEMM 8192^3 on this pool)"
        if n:
            A = torch.randn(n, n, dtype=torch.float64, device=dev)
            B = torch.randn(n, n, dtype=torch.float64, device=dev)
            (A @ B)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record(); (A @ B); b1.record(); torch.cuda.synchronize()
                best = min(best, b0.elapsed_time(b1))
            peak, peak_src = 2 * n ** 3 / best / 1e9, "cuBLAS DGEMM 8192^3 (torch.matmul float64) measured in this run"
            del A, B
        cpu = None
        if world == 1 and args.cpu_rows > 0:
            port = CpuPort()
            rate, dt, _ = port.rate(g, args.cpu_rows)
            port.close()
            # parity spot check of the timed GPU output against the oracle on the same rows
            from oracle import gp_oracle as orc
            rows = 64
            want = orc.log_posterior(sts, Xh[W + K - 1][:rows].numpy(), g["lo"], g["hi"], g["y_exp"], g["cov_exp"])
            fin = np.isfinite(want)
            cpu = {"value": rate, "unit": UNIT, "cores": port.workers, "kind": "port",
                   "sample": "%d walkers of the same workload, 128-row chunks over %d single-BLAS-thread processes, %.1f s "
                             "(NumPy/SciPy oracle port)" % (args.cpu_rows, port.workers, dt),
                   "max_abs_diff_gpu_vs_oracle": float(np.max(np.abs(lp_check[:rows][fin] - want[fin])))}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "walkers_per_gpu_per_step": N, "path": args.path,
                       "state": "hyper-parameters trained by the reference (tests/golden/c2_rbf.npz)",
                       "l2": "256 MB flush between timed steps, outside the CUDA events",
                       "collective": collective},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N * p * 8,
                    "d2h_bytes_per_step": world * N * 8, "ms_per_step": 1e3 * e2e_s / K},
            "gpu_launches": int(launches),
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / K,
            "clocks": clk.summary(),
            "roofline": {"kernel": "pc_predict_kernel (a)", "bound": "tensor", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_note": "dram read+write bytes per launch (ncu, profiles/r01_pc_predict_traffic.json); "
                                         "algorithmic HBM bytes are 8p in + 16q out = 456 B per evaluation (1.9 MB per launch): "
                                         "the rest is the 21 MB lower triangle of L^-1 streamed once into L2",
                         "hbm_achieved_GBps": (traffic or 0) / (ka_ms * 1e-3) / 1e9,
                         "hbm_peak_GBps": hbm_peak, "hbm_peak_source": hbm_src,
                         "hbm_frac": (traffic or 0) / (ka_ms * 1e-3) / 1e9 / hbm_peak,
                         "peak_source": peak_src, "ms_per_launch": ka_ms,
                         "algorithmic_flops_per_eval": flops_pc_predict(SHAPE["p"], SHAPE["n"], SHAPE["q"]),
                         "dtype": "FP64 DMMA.8x8x4 + DFMA (one shared pipe, 37.1 TFLOP/s DMMA issue peak measured)"},
            "cpu_baseline": cpu,
            "dense_path": dense,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--walkers", type=int, default=4096, help="walkers per GPU per step")
    ap.add_argument("--path", default="auto", choices=["auto", "lowrank", "dense"])
    ap.add_argument("--cpu-rows", type=int, default=None, help="rows of the CPU-baseline sample")
    ap.add_argument("--dense-steps", type=int, default=10)
    ap.add_argument("--dense-walkers", type=int, default=2048)
    ap.add_argument("--no-dgemm", dest="dgemm", action="store_false")
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        if args.cpu_rows is None:
            # bounded sample per step so that steps x rows stays near 3e5 evaluations (~1-2 min on 16 cores)
            args.cpu_rows = int(min(4096, max(256, 3e5 // max(args.steps, 1))))
        run_reference(args)
    else:
        if args.cpu_rows is None:
            args.cpu_rows = 8192
        run_ours(args)


if __name__ == "__main__":
    main()
