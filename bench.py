#!/usr/bin/env python
"""
bench.py -- log-posterior evaluations / second for BASELINE.json's config 2
(17 parameters, 500 design points, 300 observables, 20 PCs; 4096 particles per GPU per step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (Chain.log_posterior) over one fresh batch of walkers.
  value     device-timed (CUDA events on the launching stream), inputs resident in HBM, L2 flushed between
            timed steps and -- with several ranks -- a device-side barrier after the flush, both OUTSIDE the
            events, so the interval holds this step's kernels and its gather, not the other ranks' skew
  timing    per rank and step: kernels / gather (barrier) / total, min-median-max over ranks
  gather_check  (N > 1) the fused peer-store gather against an NCCL all-gather (bit for bit, both buffer
            halves) and against the oracle on rows of EVERY rank's slot
  e2e       the same metric through the host API with HOST buffers (pinned): H2D of X, kernels,
            (all-gather), D2H of lp, one synchronisation per step -- wall clock
  sustained >= 2 s of back-to-back steps without flush: throughput, SM clock, power-cap flag
  roofline, roofline_b, roofline_c, roofline_bc_fused   kernels (a), (b), (c) alone and (b)+(c) as fused on the
            dense path, at 4096 walkers, CUDA events:
            algorithmic FP64 flops or HBM bytes / time vs the cuBLAS DGEMM rate measured in this run
            (MEASURED_PEAKS.json has no FP64 entry) and vs MEASURED_PEAKS.json's hbm_gbs
  dense_path  the named-kernel contract (a) -> (b) -> (c) at 4096 walkers
  configs   compact records for BASELINE configs 1, 3, 4, 5
  fanout    (single process, several visible GPUs) Chain.log_posterior spread over 1/2/4/8 GPUs
  cpu_baseline  the oracle port (NumPy/SciPy restatement of the reference) on a bounded sample,
            on the host's cores, rank 0, N=1 only
--impl reference times the CPU restatement of the reference on the host cores (the reference itself
is pure Python that cannot travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
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
# the unmodified reference on this workload, measured once in the survey container (SURVEY.md section 6:
# Chain.log_likelihood(finite=True), N = 4096, 8 threads; its sklearn call is O(N^2) per batch)
REFERENCE_UNMODIFIED = {"value": 128.0, "unit": UNIT, "where": "survey container, 8 vCPU, SURVEY.md section 6",
                        "note": "recorded figure, not re-measured in this run: the reference is a Python tree "
                                "outside the repo and does not travel to the GPU box"}


def flops_pc_predict(p, n, q):
    """q n (3p+4) + 2 q n + q n^2 FP64 flops per evaluation (FMA = 2)"""
    return q * n * (3 * p + 4) + 2 * q * n + q * n * n


def flops_backtransform(q, m):
    """SURVEY 8(d), full symmetric output: 2 q m + 2 q m^2"""
    return 2 * q * m + 2 * q * m * m


def flops_cholesky(m):
    """m^3 / 3 + m^2"""
    return m ** 3 / 3.0 + m * m


def load_c2(keep_L=False):
    """Config-2 emulator state: hyper-parameters / alpha / PCA matrices trained by the unmodified
    reference (tests/golden/c2_rbf.npz); L_ is rebuilt from them as sklearn's fit does."""
    import gpbt_b200  # noqa: F401
    from gpbt_b200 import fixtures
    g = fixtures.load("c2_rbf")
    return g, fixtures.state_dicts(g)


def walkers(g, N, seed):
    """uniform in the box, 1 % of rows pushed out of bounds"""
    rng = np.random.default_rng(seed)
    lo, hi = g["lo"], g["hi"]
    X = rng.uniform(lo, hi, (N, len(lo)))
    rows = rng.choice(N, max(1, N // 100), replace=False)
    X[rows, rng.integers(0, len(lo), len(rows))] = hi.max() + 1.0
    return X


# ---- clocks: NVML polled by a CHILD PROCESS (no GIL, no driver calls from the timed process) -------
_SAMPLER_SRC = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
key, period = sys.argv[1], float(sys.argv[2])
h = nv.nvmlDeviceGetHandleByPciBusId(key.encode()) if ":" in key else nv.nvmlDeviceGetHandleByIndex(int(key))
sys.stdout.write("max %d\n" % nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)); sys.stdout.flush()
while True:
    t = time.time()
    try:
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        pw = nv.nvmlDeviceGetPowerUsage(h)
    except Exception:
        break
    sys.stdout.write("%.6f %d %d %d\n" % (t, sm, rs, pw)); sys.stdout.flush()
    time.sleep(period)
"""
_REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}


class ClockSampler:
    """SM clock, throttle reasons and power of one GPU, polled every `period` s for the life of the
    object; summary(t0, t1) reports the samples that fall into a wall-clock window."""

    def __init__(self, torch, index, period=0.002, mode="process"):
        self.mode, self.lines, self.max_mhz = mode, [], None
        self.proc = self.thread = None
        if mode == "off":
            return
        key = self._device_key(torch, index)
        if mode == "process":
            try:
                self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, key, str(period)],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._drain, daemon=True)
                self.thread.start()
            except OSError:
                self.proc = None

    @staticmethod
    def _device_key(torch, index):
        try:
            bus = torch.cuda.get_device_properties(index).pci_bus_id
            dom = getattr(torch.cuda.get_device_properties(index), "pci_domain_id", 0)
            dev = torch.cuda.get_device_properties(index).pci_device_id
            return "%08X:%02X:%02X.0" % (dom, bus, dev)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    return str(int(vis.split(",")[index]))
                except (ValueError, IndexError):
                    pass
            return str(index)

    def _drain(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def close(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.thread.join(timeout=2)
            self.proc = None

    def summary(self, t0, t1):
        sm, power, reasons = [], [], 0
        for line in list(self.lines):
            f = line.split()
            if f and f[0] == "max":
                self.max_mhz = float(f[1])
            elif len(f) == 4 and t0 <= float(f[0]) <= t1:
                sm.append(float(f[1]))
                reasons |= int(f[2])
                power.append(float(f[3]) / 1e3)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(np.min(sm)), "sm_max_mhz": self.max_mhz,
                "reasons": [k for k, bit in _REASONS.items() if reasons & bit], "samples": len(sm),
                "power_w_max": float(np.max(power))}


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
        "config": {"workload": WORKLOAD, "walkers_per_gpu_per_step": rows,
                   "note": "CPU restatement of the reference: the reference itself is pure Python outside the repo and "
                           "cannot travel to the GPU box; its own sklearn call is O(N^2) per batch and slower than this port"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": port.workers, "kind": "port", "sample": sample},
        "reference_unmodified": REFERENCE_UNMODIFIED,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def _stats(a):
    a = np.asarray(a, dtype=np.float64)
    return {"min": float(a.min()), "median": float(np.median(a)), "max": float(a.max())}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gpbt_b200  # noqa: F401
    from gpbt_b200 import _lib, fixtures
    from gpbt_b200.device import DeviceChain, DeviceEmulator
    from gpbt_b200.dist import ShardedEvaluator

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
    clk = ClockSampler(torch, local, mode=args.clock_sampler)      # running before any warm-up

    g = fixtures.load("c2_rbf")
    states, sts = fixtures.emulator_states(g)
    y_exp = g["y_exp"].reshape(-1)
    chain = DeviceChain(states, g["lo"], g["hi"], y_exp, g["cov_exp"], devices=[local])
    N, p = args.walkers, len(g["lo"])
    K, W = args.steps, args.warmup
    nb = K + W
    # fresh walkers every step; each rank has its own rows (weak scaling)
    seed_of = lambda r, i: 1000 * r + i   # noqa: E731
    Xh = [torch.from_numpy(walkers(g, N, seed_of(rank, i))).pin_memory() for i in range(nb)]
    Xd = [x.to(dev) for x in Xh]
    ev = ShardedEvaluator(lambda X: chain.log_target_device(X, -np.inf, path=args.path), dev)
    # multi-GPU: the all-gather of lp is fused into the last kernel (NVLink peer stores + one device
    # barrier); NCCL all-gather is the fallback (--collective nccl, or symmetric memory unavailable)
    collective = "none"
    pg = None
    align = torch.zeros(1, device=dev)

    def gather_fn(X, mid_event=None):
        lp = ev.eval_fn(X)
        if mid_event is not None:
            mid_event.record()
        return ev.gather_shards(lp, X.shape[0])

    def pre_step_align():
        if world > 1:
            dist.all_reduce(align)     # device-side: the stream waits for every rank, the host does not

    if world > 1:
        collective = "nccl all-gather of lp (8 B/walker)"
        if args.collective == "fused":
            try:
                from gpbt_b200.dist import PeerGather
                pg = PeerGather(N, dev)
                gather_fn = lambda X, mid_event=None: pg.evaluate(chain, X, -np.inf, path=args.path, mid_event=mid_event)   # noqa: E731
                pre_step_align = lambda: pg.hdl.barrier(channel=1)   # noqa: E731
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

    def gather_obj(obj):
        if world == 1:
            return [obj]
        out = [None] * world if rank == 0 else None
        dist.gather_object(obj, out, dst=0)
        return out

    # ---- device-resident leg -------------------------------------------------------------
    for i in range(W):
        flush.fill_(i & 0xff)
        pre_step_align()
        gather_fn(Xd[i])
    barrier()
    launches0 = _lib.lib.gpbt_launch_count()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    evm = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    t_region0 = time.time()
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush.fill_(i & 0xff)
        pre_step_align()              # ranks leave the flush together; outside the timed interval
        ev0[i].record()
        out = gather_fn(Xd[W + i], mid_event=evm[i])
        ev1[i].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    t_region1 = time.time()
    launches = _lib.lib.gpbt_launch_count() - launches0
    kern_ms = [a.elapsed_time(b) for a, b in zip(ev0, evm)]
    gath_ms = [a.elapsed_time(b) for a, b in zip(evm, ev1)]
    tot_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    dev_ms = maxr(sum(tot_ms))
    med_ms = maxr(float(np.median(tot_ms)))
    value = world * N * K / (dev_ms * 1e-3)
    lp_check = out[:N].cpu().numpy() if world == 1 else None
    per_rank = gather_obj({"rank": rank, "kernels_ms": _stats(kern_ms), "gather_ms": _stats(gath_ms),
                           "total_ms": _stats(tot_ms), "sum_total_ms": float(sum(tot_ms)),
                           "slowest_step": int(np.argmax(tot_ms))})
    clocks = clk.summary(t_region0, t_region1)

    # ---- gather check (N > 1): fused buffer vs NCCL all-gather, both halves; oracle on every slot ----
    gather_check = None
    if world > 1:
        worst, rows_checked = 0.0, 0
        mism = 0
        orc_err = None
        for rep in range(2):                       # two consecutive calls = both halves of the buffer
            i = W + K - 1 - rep
            got = gather_fn(Xd[i]).clone()
            lp_local = chain.log_target_device(Xd[i], -np.inf, path=args.path)
            want = torch.empty(world * N, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(want, lp_local.contiguous())
            same = (got == want) | (torch.isinf(got) & torch.isinf(want) & (got.sign() == want.sign()))
            mism += int((~same).sum().item())
            fin = torch.isfinite(want) & torch.isfinite(got)
            if fin.any():
                worst = max(worst, float((got[fin] - want[fin]).abs().max().item()))
            if rank == 0 and rep == 0:
                from oracle import gp_oracle as orc          # checker only, after the timed region
                got_h = got.cpu().numpy()
                orc_err = 0.0
                for r in range(world):
                    Xr = walkers(g, N, seed_of(r, i))[:8]
                    ref = orc.log_posterior(sts, Xr, g["lo"], g["hi"], g["y_exp"], g["cov_exp"])
                    slot = got_h[r * N:r * N + 8]
                    okf = np.isfinite(ref)
                    if not np.array_equal(np.isfinite(slot), okf):
                        orc_err = float("inf")
                    elif okf.any():
                        orc_err = max(orc_err, float(np.max(np.abs(slot[okf] - ref[okf]))))
                    rows_checked += 8
        mism = int(maxr(float(mism)))
        worst = maxr(worst)
        gather_check = {"gather_max_abs_diff": worst, "mismatching_entries": mism,
                        "what": "fused peer-store gather vs NCCL all_gather_into_tensor of the ranks' own lp, "
                                "both buffer halves, all %d entries on every rank" % (world * N),
                        "oracle_max_abs_diff_all_slots": orc_err, "oracle_rows": rows_checked}
        if mism != 0 or (orc_err is not None and not orc_err <= 1e-8):
            raise SystemExit("bench.py: gathered log-posteriors are wrong: %r" % (gather_check,))

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
        e2e_step(W + i)
    barrier()
    e2e_s = maxr(time.perf_counter() - t0)
    e2e_value = world * N * K / e2e_s

    # ---- sustained leg: >= 2 s of back-to-back steps, no flush ----------------------------------
    sustained = None
    if args.sustained_s > 0:
        reps = max(K, int(args.sustained_s / (med_ms * 1e-3)) + 1)
        reps = int(maxr(float(reps)))
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts0 = time.time()
        s0.record()
        for i in range(reps):
            gather_fn(Xd[i % nb])
        s1.record()
        barrier()
        ts1 = time.time()
        s_ms = maxr(s0.elapsed_time(s1))
        sc = clk.summary(ts0, ts1)
        sustained = {"value": world * N * reps / (s_ms * 1e-3), "unit": UNIT, "seconds": s_ms * 1e-3, "steps": reps,
                     "ms_per_step": s_ms / reps, "l2": "no flush: back-to-back sampler steps",
                     "sm_mhz_median": sc.get("sm_mhz"), "sm_mhz_min": sc.get("sm_min_mhz"),
                     "power_w_max": sc.get("power_w_max"), "sw_power_cap": "sw_power_cap" in sc.get("reasons", []),
                     "reasons": sc.get("reasons")}

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "walkers_per_gpu_per_step": N, "path": args.path,
                       "state": "hyper-parameters trained by the reference (tests/golden/c2_rbf.npz)",
                       "l2": "256 MB flush between timed steps, outside the CUDA events",
                       "collective": collective, "clock_sampler": args.clock_sampler},
            "value_from_median_step": world * N / (med_ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N * p * 8,
                    "d2h_bytes_per_step": world * N * 8, "ms_per_step": 1e3 * e2e_s / K,
                    "note": "no L2 flush in this leg (back-to-back host calls): kernel (a) finds its 21 MB of L^-1 in L2"},
            "gpu_launches": int(launches),
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / K,
            "clocks": clocks,
            "timing": {"per_rank": per_rank,
                       "kernels_ms_median_over_ranks": _stats([r["kernels_ms"]["median"] for r in per_rank]),
                       "gather_ms_median_over_ranks": _stats([r["gather_ms"]["median"] for r in per_rank]),
                       "total_ms_median_over_ranks": _stats([r["total_ms"]["median"] for r in per_rank]),
                       "note": "CUDA events per rank and step: start -> kernels done -> gather (device barrier) done"},
            "gather_check": gather_check,
            "sustained": sustained,
        }
        line.update(single_gpu_records(args, torch, dev, chain, states, sts, g, Xh, Xd, lp_check, world, W, K))
    clk.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


def _event_ms(torch, fn, reps, warm=2, per_call=True):
    """CUDA-event time of one fn().  per_call: median over reps, one event pair per call (a host hiccup while a
    call's 40-odd launches are being enqueued inflates that call only, not the figure); else the mean over reps
    back-to-back calls between one event pair (single-kernel calls: the average launch duration)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if not per_call:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    evs = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs]))


def single_gpu_records(args, torch, dev, chain, states, sts, g, Xh, Xd, lp_check, world, W, K):
    """rank 0 only, after the timed legs: per-kernel rooflines, dense path, config records, CPU baseline"""
    import ctypes as C
    from gpbt_b200 import _lib, synthetic
    from gpbt_b200.device import DeviceChain, DeviceEmulator, _stream_ptr
    from gpbt_b200.state import EmulatorState
    N, nb = args.walkers, len(Xd)
    p, n, m, q = SHAPE["p"], SHAPE["n"], SHAPE["m"], SHAPE["q"]
    rec = {}
    de = DeviceEmulator(states[0])
    st = states[0]
    # HBM denominator: the driver's measured copy rate, else the profiling guide's fallback
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm_peak, hbm_src = float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except (OSError, KeyError, ValueError):
        pass

    # ---- kernel (a) alone ------------------------------------------------------------------
    cnt = [0]

    def run_a():
        cnt[0] += 1
        return de.pc_predict_device(Xd[cnt[0] % nb])
    ka_ms = _event_ms(torch, run_a, max(K, 10), warm=3, per_call=False)
    fl_a = flops_pc_predict(p, n, q)

    # ---- kernel (b) alone: covariance materialised, full 4096 walkers ---------------------------
    zm, zv = de.pc_predict_device(Xd[0])
    mean = torch.empty((N, m), dtype=torch.float64, device=dev)
    cov = torch.empty((N, m, m), dtype=torch.float64, device=dev)

    def run_b():
        _lib.check(_lib.lib.gpbt_backtransform(st.handle(), zm.data_ptr(), zv.data_ptr(), q, mean.data_ptr(), m,
                                               cov.data_ptr(), m, 0, N, _stream_ptr(torch)))
    kb_ms = _event_ms(torch, run_b, 5)
    bytes_b = 8 * (m + m * m)

    # ---- kernel (c) alone: batched Cholesky log-likelihood of those covariances (+ cov_exp) --------
    cov_exp_d = torch.from_numpy(np.ascontiguousarray(g["cov_exp"])).to(dev)
    y_d = torch.from_numpy(np.ascontiguousarray(g["y_exp"].reshape(-1))).to(dev)
    lp_c = torch.empty(N, dtype=torch.float64, device=dev)
    kc = []
    for _ in range(4):
        run_b()                                   # (c) factorises in place: fresh covariances every time
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        _lib.check(_lib.lib.gpbt_mvn_loglike(mean.data_ptr(), y_d.data_ptr(), cov.data_ptr(), cov_exp_d.data_ptr(),
                                             lp_c.data_ptr(), None, float("-inf"), N, m, _stream_ptr(torch)))
        c1.record()
        torch.cuda.synchronize()
        kc.append(c0.elapsed_time(c1))
    kc_ms = float(np.median(kc[1:]))
    del cov, mean

    # ---- dense path: the named-kernel contract (a) -> (b) -> (c), reported separately; timed before the
    # DGEMM below, whose power draw lowers the clocks of whatever runs right after it ----------------
    dense = None
    if args.dense_steps > 0:
        Nd = args.dense_walkers
        dcnt = [0]

        def run_dense():
            dcnt[0] += 1
            return chain.log_target_device(Xd[dcnt[0] % nb][:Nd], -np.inf, path="dense")
        dms = _event_ms(torch, run_dense, args.dense_steps, warm=3)
        lpd = chain.log_target_device(Xd[0][:Nd], -np.inf, path="dense")
        ref = chain.log_target_device(Xd[0][:Nd], -np.inf, path="lowrank")
        fin = torch.isfinite(ref)
        fl_all = fl_a + flops_backtransform(q, m) + flops_cholesky(m)
        dense = {"value": Nd / (dms * 1e-3), "unit": UNIT, "walkers": Nd, "ms_per_step": dms,
                 "max_abs_diff_vs_lowrank": float((lpd[fin] - ref[fin]).abs().max().item()),
                 "algorithmic_flops_per_eval": fl_all, "achieved_TFLOPs": fl_all * Nd / (dms * 1e-3) / 1e12}

    # ---- other BASELINE configurations (compact) ----------------------------------------------------
    configs = config_records(args, torch, dev, chain, states, g) if (args.configs and world == 1) else None

    # ---- FP64 roofline denominator: cuBLAS DGEMM, measured here --------------------------------
    nn = 8192 if args.dgemm else 0
    peak, peak_src = 35.45, "profiles/r01_dgemm_peak.json (cuBLAS DGEMM 8192^3 on this pool)"
    if nn:
        A = torch.randn(nn, nn, dtype=torch.float64, device=dev)
        B = torch.randn(nn, nn, dtype=torch.float64, device=dev)
        (A @ B)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record(); (A @ B); b1.record(); torch.cuda.synchronize()
            best = min(best, b0.elapsed_time(b1))
        peak, peak_src = 2 * nn ** 3 / best / 1e9, "cuBLAS DGEMM 8192^3 (torch.matmul float64) measured in this run"
        del A, B
    if dense is not None:
        dense["frac_of_dgemm"] = dense["achieved_TFLOPs"] / peak

    # DRAM traffic of the dominant kernel per launch, from the committed ncu capture
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_pc_predict_traffic.json")) as fh:
            tj = json.load(fh)
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except (OSError, KeyError, ValueError):
        pass
    ach_a = fl_a * N / (ka_ms * 1e-3) / 1e12
    rec["roofline"] = {
        "kernel": "pc_predict_kernel (a)", "bound": "tensor", "achieved": ach_a, "peak": peak,
        "unit": "TFLOP/s", "frac": ach_a / peak, "traffic": traffic,
        "traffic_note": "dram read+write bytes per launch (ncu, profiles/r01_pc_predict_traffic.json); "
                        "algorithmic HBM bytes are 8p in + 16q out = 456 B per evaluation (1.9 MB per launch): "
                        "the rest is the 21 MB lower triangle of L^-1 streamed once into L2",
        "hbm_achieved_GBps": (traffic or 0) / (ka_ms * 1e-3) / 1e9,
        "hbm_peak_GBps": hbm_peak, "hbm_peak_source": hbm_src,
        "hbm_frac": (traffic or 0) / (ka_ms * 1e-3) / 1e9 / hbm_peak,
        "peak_source": peak_src, "ms_per_launch": ka_ms, "walkers_per_launch": N,
        "algorithmic_flops_per_eval": fl_a,
        "dtype": "FP64 DMMA.8x8x4 + DFMA (one shared pipe, 37.1 TFLOP/s DMMA issue peak measured)"}
    ach_b = bytes_b * N / (kb_ms * 1e-3) / 1e9
    rec["roofline_b"] = {
        "kernel": "backtransform_mean_kernel + backtransform_cov_kernel (b), covariance materialised in HBM",
        "bound": "hbm", "achieved": ach_b, "peak": hbm_peak, "unit": "GB/s", "frac": ach_b / hbm_peak, "traffic": None,
        "peak_source": hbm_src, "ms_per_launch": kb_ms, "walkers_per_launch": N,
        "algorithmic_bytes_per_eval": bytes_b, "algorithmic_flops_per_eval": flops_backtransform(q, m),
        "achieved_TFLOPs": flops_backtransform(q, m) * N / (kb_ms * 1e-3) / 1e12}
    ach_c = flops_cholesky(m) * N / (kc_ms * 1e-3) / 1e12
    rec["roofline_c"] = {
        "kernel": "gpbt_mvn_loglike (c): batched Cholesky + solve + log-det of materialised covariances (+ cov_exp)",
        "bound": "tensor", "achieved": ach_c, "peak": peak, "unit": "TFLOP/s", "frac": ach_c / peak, "traffic": None,
        "peak_source": peak_src, "ms_per_launch": kc_ms, "walkers_per_launch": N,
        "algorithmic_flops_per_eval": flops_cholesky(m), "algorithmic_bytes_per_eval": 8 * (m * m + m),
        "hbm_GBps_algorithmic": 8 * (m * m + m) * N / (kc_ms * 1e-3) / 1e9}
    if dense is not None and dense["walkers"] == N:
        # (b) + (c) as they run on the dense path: fused (the covariance is generated inside the Cholesky
        # launches); their time = dense step - kernel (a) alone (the mean kernel, ~1 %, stays in)
        fl_bc = 2 * q * m + q * m * (m + 1) + flops_cholesky(m)      # (b) lower triangle + (c)
        bc_ms = dense["ms_per_step"] - ka_ms
        ach_bc = fl_bc * N / (bc_ms * 1e-3) / 1e12
        rec["roofline_bc_fused"] = {
            "kernel": "chol_fused_panel_kernel + chol_fused_factor_kernel: kernels (b) + (c) fused, dense path",
            "bound": "tensor", "achieved": ach_bc, "peak": peak, "unit": "TFLOP/s", "frac": ach_bc / peak, "traffic": None,
            "peak_source": peak_src, "ms_per_launch": bc_ms, "walkers_per_launch": N,
            "how": "dense-path step minus kernel (a) timed alone, both CUDA events in this run",
            "algorithmic_flops_per_eval": fl_bc,
            "note": "q m (m + 1) + 2 q m for the lower triangle of (b), m^3 / 3 + m^2 for (c); nothing of size N m^2 "
                    "is written: the factor (lower blocks only, 332 KB per walker) is the only per-walker HBM traffic"}
    rec["dense_path"] = dense
    rec["configs"] = configs

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
               "max_abs_diff_gpu_vs_oracle": float(np.max(np.abs(lp_check[:rows][fin] - want[fin]))),
               "reference_unmodified": REFERENCE_UNMODIFIED}
    rec["cpu_baseline"] = cpu
    return rec


def _host_rate(fn, X, reps):
    fn(X)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        lp = fn(X)
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    return len(X) / dt, dt, lp


def config_records(args, torch, dev, chain2, states2, g2):
    """BASELINE configs 1, 3, 4, 5 on this GPU (config 2 is the headline), plus -- when this single process
    sees several GPUs -- Chain.log_posterior's fan-out over them."""
    from gpbt_b200 import _lib, fixtures, synthetic
    from gpbt_b200.device import DeviceChain, DeviceEmulator
    from gpbt_b200.state import EmulatorState
    out = {}
    # C1: p5 n100 m50 q10, emcee 128 walkers -> host calls of 64 rows (latency bound)
    g1 = fixtures.load("c1_rbf")
    st1, _ = fixtures.emulator_states(g1)
    ch1 = DeviceChain(st1, g1["lo"], g1["hi"], g1["y_exp"].reshape(-1), g1["cov_exp"], devices=[dev.index])
    X1 = np.ascontiguousarray(g1["X"][:64])
    rate, dt, lp = _host_rate(lambda X: ch1.log_target(X, -np.inf), X1, 300)
    fin = np.isfinite(g1["lp_posterior"][:64])
    out["C1"] = {"shape": "p5 n100 m50 q10", "rows_per_call": 64, "us_per_call": dt * 1e6, "evals_per_s": rate,
                 "what": "Chain.log_posterior host call of one emcee half-ensemble (64 of 128 walkers), median of 300",
                 "max_abs_diff_vs_reference_golden": float(np.max(np.abs(lp[fin] - g1["lp_posterior"][:64][fin])))}
    ch1.release()
    # C4: C2 state, 2^20 walkers per call, full (non-diagonal) experimental covariance, host API
    cov_sys = g2["cov_exp"] + synthetic.systematic_cov(g2["cov_exp"].shape[0])
    ch4 = DeviceChain(states2, g2["lo"], g2["hi"], g2["y_exp"].reshape(-1), cov_sys, devices=[dev.index])
    X4 = walkers(g2, args.c4_rows, 6)
    rate, dt, lp = _host_rate(lambda X: ch4.log_target(X, -np.inf), X4, 2)
    lpg = ch4.log_target(g2["X"], -np.inf)
    fin = np.isfinite(g2["lp_posterior_sys"])
    out["C4"] = {"shape": "C2 state, full Sigma_exp", "rows_per_call": len(X4), "evals_per_s": rate, "ms_per_call": dt * 1e3,
                 "n_gpus": 1, "what": "Chain.log_posterior host call (pageable X in, lp out)",
                 "max_abs_diff_vs_reference_golden": float(np.max(np.abs(lpg[fin] - g2["lp_posterior_sys"][fin])))}
    # C5: posterior-predictive sweep, Emulator.predict over LHD points (device resident chunks)
    de = DeviceEmulator(states2[0])
    Xd = torch.from_numpy(walkers(g2, 1 << 17, 8)).to(dev)
    ms = _event_ms(torch, lambda: de.predict_diag_device(Xd), 4)
    out["C5_diag"] = {"what": "Emulator.predict mean + diag(cov), 2^17-row device chunks", "points_per_s": Xd.shape[0] / (ms * 1e-3)}
    rows = args.c5_cov_rows
    ms = _event_ms(torch, lambda: de.predict_device(Xd[:rows], True), 4)
    out["C5_cov"] = {"what": "Emulator.predict(return_cov=True), covariance materialised in HBM", "rows_per_chunk": rows,
                     "points_per_s": rows / (ms * 1e-3), "cov_write_GBps": rows * 300 * 300 * 8 / (ms * 1e-3) / 1e9}
    del Xd
    # C3: surmise-PCGP-shaped emulator (EmulatorBAND path, kernel kind 2), p15 n1000 m300 q20, 8192 chains
    if args.c3:
        info = synthetic.pcgp_fitinfo(15, 1000, 300, 20)
        stb = EmulatorState.from_pcgp_fitinfo(info)
        lo, hi = synthetic.box(15)
        yb = info["offset"] + 0.3 * info["scale"]
        chb = DeviceChain([stb], lo, hi, yb, np.diag((0.03 * np.abs(yb)) ** 2), devices=[dev.index])
        X3 = synthetic.walkers(15, 8192, seed=9)
        rate, dt, lp = _host_rate(lambda X: chb.log_target(X, -np.inf), X3, 3)
        out["C3"] = {"shape": "p15 n1000 m300 q20, surmise-PCGP-shaped fit (parity unpinned: surmise absent)",
                     "rows_per_call": 8192, "evals_per_s": rate, "ms_per_call": dt * 1e3,
                     "what": "Chain.log_posterior host call, all PTLMC chains at once"}
        # one PTLMC iteration over those 8192 chains (7168 tempered + 1024 at T = 1): the host loop of
        # gpbt_b200.ptlmc (NumPy proposal / accept, one GPU call, exchange sweeps in the C helper) and the
        # device-resident loop (gpbt_ptlmc_*)
        from gpbt_b200 import ptlmc
        n_all, n_hot, iters = 8192, 7168, 10
        temps = ptlmc.temperature_ladder(n_hot, n_all - n_hot, 100.0)
        root = np.diag(0.02 * (hi - lo))
        theta = X3.copy()
        f = chb.log_target(theta, -np.inf).reshape(-1, 1) / temps
        np.random.seed(0)
        t0 = time.perf_counter()
        for _ in range(iters):
            prop = theta + np.sqrt(2) * temps ** (1 / 3) * (np.random.normal(0, 1, theta.shape) @ root)
            fp = chb.log_target(prop, -np.inf).reshape(-1, 1) / temps
            take = np.where(np.log(np.random.uniform(size=n_all)) < np.squeeze(fp - f))[0]
            theta[take], f[take] = prop[take], fp[take]
            flat = f * temps
            order = ptlmc.temp_exchange(flat, temps, iters=5)
            f, theta = flat[order] / temps, theta[order]
        ms_host = (time.perf_counter() - t0) / iters * 1e3
        pt = ptlmc.DevicePTLMC(chb, temps, root, n_hot, seed=1)
        pt.set_state(X3, tau=-1.0)
        pt.run(2 * iters, iters, n_steps=iters)           # (warm-up)
        t0 = time.perf_counter()
        pt.run(2 * iters, iters, n_steps=2 * iters)
        ms_dev = (time.perf_counter() - t0) / (2 * iters) * 1e3
        pt.close()
        out["C3"]["ptlmc_iteration"] = {"chains": n_all, "at_T1": n_all - n_hot, "ms_host_loop": ms_host,
                                        "ms_device_loop": ms_dev,
                                        "what": "proposal + log-posterior of all chains + tempered accept + 5 exchange sweeps"}
        chb.release()
        stb.release()
    # fan-out: one process, several GPUs, through the same host call
    ndev = _lib.lib.gpbt_device_count()
    if ndev > 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1 and args.fanout:
        devs = [dev.index] + [d for d in range(ndev) if d != dev.index]
        chf = DeviceChain(states2, g2["lo"], g2["hi"], g2["y_exp"].reshape(-1), g2["cov_exp"], devices=devs)
        ch4f = DeviceChain(states2, g2["lo"], g2["hi"], g2["y_exp"].reshape(-1), cov_sys, devices=devs)
        X2 = walkers(g2, 4096, 5)
        ref2 = chf.log_target(X2, -np.inf, max_devices=1)
        fan = {"devices": ndev, "C2_4096_strong": [], "C4_strong": []}
        counts = [c for c in (1, 2, 4, 8) if c <= ndev]
        for c in counts:
            rate, dt, lp = _host_rate(lambda X: chf.log_target(X, -np.inf, max_devices=c), X2, 30)
            okf = np.isfinite(ref2)
            fan["C2_4096_strong"].append({"gpus": c, "evals_per_s": rate, "ms_per_call": dt * 1e3,
                                          "max_abs_diff_vs_1gpu": float(np.max(np.abs(lp[okf] - ref2[okf]))) if
                                          np.array_equal(np.isfinite(lp), okf) else float("inf")})
        ref4 = None
        for c in counts:
            rate, dt, lp = _host_rate(lambda X: ch4f.log_target(X, -np.inf, max_devices=c), X4, 2)
            if ref4 is None:
                ref4 = lp
            okf = np.isfinite(ref4)
            fan["C4_strong"].append({"gpus": c, "rows_per_call": len(X4), "evals_per_s": rate, "ms_per_call": dt * 1e3,
                                     "max_abs_diff_vs_1gpu": float(np.max(np.abs(lp[okf] - ref4[okf]))) if
                                     np.array_equal(np.isfinite(lp), okf) else float("inf")})
        # crossover: smallest batch for which two GPUs beat one
        cross = None
        for rows in (512, 1024, 2048, 4096, 8192):
            Xc = walkers(g2, rows, 11)
            r1 = _host_rate(lambda X: chf.log_target(X, -np.inf, max_devices=1), Xc, 20)[0]
            r2 = _host_rate(lambda X: chf.log_target(X, -np.inf, max_devices=2), Xc, 20)[0]
            fan.setdefault("crossover_probe", []).append({"rows": rows, "evals_per_s_1gpu": r1, "evals_per_s_2gpu": r2})
            if cross is None and r2 > r1:
                cross = rows
        fan["two_gpus_win_from_rows"] = cross
        out["fanout"] = fan
        chf.release()
        ch4f.release()
    ch4.release()
    return out


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
    ap.add_argument("--dense-walkers", type=int, default=4096)
    ap.add_argument("--sustained-s", type=float, default=2.0)
    ap.add_argument("--no-dgemm", dest="dgemm", action="store_false")
    ap.add_argument("--no-configs", dest="configs", action="store_false")
    ap.add_argument("--no-c3", dest="c3", action="store_false")
    ap.add_argument("--no-fanout", dest="fanout", action="store_false")
    ap.add_argument("--c4-rows", type=int, default=1 << 20)
    ap.add_argument("--c5-cov-rows", type=int, default=8192)
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"])
    ap.add_argument("--clock-sampler", default="process", choices=["process", "off"])
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
