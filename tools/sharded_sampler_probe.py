"""torchrun --nproc-per-node N tools/sharded_sampler_probe.py [walkers] [steps]
Every rank runs the sharded ensemble sampler (replicated walkers, proposals evaluated in slices, fused
gather of the log-posteriors) and the single-GPU sampler with the same seed on the config-2 chain: the
chains must be identical on every rank and equal to the single-GPU chain; prints steps/s for both."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gpbt_b200  # noqa: E402,F401
from gpbt_b200.device import DeviceChain  # noqa: E402
from gpbt_b200.sampler import DeviceEnsembleSampler, ShardedEnsembleSampler  # noqa: E402
from gpbt_b200.state import EmulatorState  # noqa: E402
from tests import goldens  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 1001
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12

g = goldens.load("c2_rbf")
sts = goldens.oracle_states(g)
states = [EmulatorState.from_arrays(s["kind"], s["Xtr"], s["ell"], s["c"], s["sn"], s["alpha"], s["mu"],
                                    s["scale"], s.get("A"), s.get("Ctrunc"), L=s["L"], keep_L=False) for s in sts]
chain = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
rng = np.random.default_rng(5)
x0 = 0.5 * (g["lo"] + g["hi"]) + 0.25 * (g["hi"] - g["lo"]) * rng.uniform(-1, 1, (nw, len(g["lo"])))

sh = ShardedEnsembleSampler(nw, x0.shape[1], chain, seed=77)
sh.set_state(x0)
sh.advance(2)                                   # warm-up (workspaces, symmetric buffers)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
sh.advance(steps)
dist.barrier(); torch.cuda.synchronize()
dt_sh = time.perf_counter() - t0
got = sh.get_chain()
got_lp = sh.get_log_prob()

single = DeviceEnsembleSampler(nw, x0.shape[1], chain, seed=77)
single.set_state(x0)
single.advance(2)
torch.cuda.synchronize()
t0 = time.perf_counter()
single.advance(steps)
dt_1 = time.perf_counter() - t0
want = single.get_chain()
want_lp = single.get_log_prob()

same_pos = bool(np.array_equal(got, want))
fin = np.isfinite(want_lp)
max_lp = float(np.max(np.abs(got_lp[fin] - want_lp[fin])))
# identical on every rank: compare a checksum through NCCL
chk = torch.tensor([float(np.sum(got)), float(np.sum(got_lp[np.isfinite(got_lp)]))], dtype=torch.float64, device="cuda")
lo_, hi_ = chk.clone(), chk.clone()
dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
same_ranks = bool(torch.equal(lo_, hi_))
ok = same_pos and same_ranks and max_lp <= 1e-9 and got.shape == (steps + 2, nw, x0.shape[1])
if rank == 0:
    print({"world": world, "walkers": nw, "steps": steps, "sharded_steps_per_s": steps / dt_sh,
           "single_gpu_steps_per_s": steps / dt_1, "speedup": dt_1 / dt_sh, "same_chain": same_pos,
           "same_on_all_ranks": same_ranks, "max_lp_diff": max_lp,
           "acceptance": float(sh.acceptance_fraction.mean())})
    print("SHARDED_SAMPLER_PASS" if ok else "SHARDED_SAMPLER_FAIL")
sh.close(); single.close(); chain.release()

# Chain.run_mcmc under torchrun picks the sharded sampler by itself; the chain file is written by rank 0
import pickle  # noqa: E402
import tempfile  # noqa: E402

from gpbt_b200 import synthetic  # noqa: E402
from gpbt_b200.mcmc import Chain  # noqa: E402
from tests.helpers import product_states  # noqa: E402

g1 = goldens.load("c1_rbf")
tmp = tempfile.mkdtemp(prefix="gpbt_rank%d_" % rank)
paths = synthetic.write_fixture(tmp, p=5, n=8, m=50)
os.makedirs(os.path.join(tmp, "mcmc"), exist_ok=True)
ch = Chain(mcmc_path=os.path.join(tmp, "mcmc", "chain.pkl"), expdata_path=paths["exp"], model_parafile=paths["par"])
ch.emuList = product_states(g1)[0]
np.random.seed(3)
ch.run_mcmc(nsteps=12, nburnsteps=8, nwalkers=20, nthin=2, seed=9, sampler="device")
chk = torch.tensor([float(np.sum(ch.chain))], dtype=torch.float64, device="cuda")
lo_, hi_ = chk.clone(), chk.clone()
dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
ok2 = bool(torch.equal(lo_, hi_)) and ch.chain.shape == (20, 6, 5) and (os.path.exists(ch.mcmc_path) == (rank == 0))
if rank == 0:
    with open(ch.mcmc_path, "rb") as fh:
        ok2 = ok2 and pickle.load(fh)["chain"].shape == (20, 6, 5)
    print("RUN_MCMC_SHARDED_PASS" if ok2 else "RUN_MCMC_SHARDED_FAIL")
ok = ok and ok2
dist.destroy_process_group()
sys.exit(0 if ok else 1)
