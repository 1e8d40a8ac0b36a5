"""PeerGather (fused all-gather over peer-mapped buffers) on 2+ GPUs, checked against an NCCL
all-gather of the same values.  Launched by tests/test_multi_gpu.py under torchrun."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import gpbt_b200  # noqa
from gpbt_b200.device import DeviceChain
from gpbt_b200.dist import PeerGather
from tests import goldens
from tests.helpers import product_states

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = goldens.load("c1_rbf")
states, _ = product_states(g)
ch = DeviceChain(states, g["lo"], g["hi"], g["y_exp"].reshape(-1), g["cov_exp"])
N = 64
X = torch.from_numpy(np.ascontiguousarray(g["X"][rank * N:(rank + 1) * N] if world * N <= len(g["X"]) else g["X"][:N])).to(dev)
want_local = ch.log_target_device(X, -np.inf)
torch.cuda.synchronize()
print(rank, "local ok", flush=True)
pg = PeerGather(N, dev)
print(rank, "ptrs", [hex(p) for p in pg.ptrs], "buf", hex(pg.buf.data_ptr()), pg.buf.numel(), flush=True)
# write through raw pointers with a trivial torch-free path first: our scatter kernel on the dense path
for path in ("dense", "lowrank"):
    want_local = ch.log_target_device(X, -np.inf, path=path)
    out = pg.evaluate(ch, X, -np.inf, path=path)
    torch.cuda.synchronize()
    mine = out[rank * N:(rank + 1) * N]
    fin = torch.isfinite(want_local)
    print(rank, path, "own slice max diff", float((mine[fin] - want_local[fin]).abs().max()), flush=True)
    gathered = [torch.empty_like(want_local) for _ in range(world)]
    dist.all_gather(gathered, want_local)
    ref = torch.cat(gathered)
    fin = torch.isfinite(ref)
    err = float((out[fin] - ref[fin]).abs().max())
    print(rank, path, "gathered max diff", err, flush=True)
    assert err == 0.0 and torch.equal(torch.isinf(out), torch.isinf(ref)), (path, err)
# several calls in a row: the two halves of the symmetric buffer alternate
for i in range(5):
    out = pg.evaluate(ch, X, -np.inf)
    assert float((out[fin] - ref[fin]).abs().max()) == 0.0
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("PEER_GATHER_PASS", flush=True)
dist.destroy_process_group()
