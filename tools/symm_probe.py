"""Probe: does torch symmetric memory (peer-mapped buffers + device barrier) work on this box?"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    t = symm_mem.empty(world * 8, dtype=torch.float64, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok", [hex(p) for p in hdl.buffer_ptrs][:4], flush=True)
    t.fill_(-1.0)
    hdl.barrier(channel=0)
    for r in range(world):
        peer = hdl.get_buffer(r, (world * 8,), torch.float64)
        peer[rank * 8:(rank + 1) * 8] = float(rank + 1)
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    print(rank, "gathered", t.view(world, 8)[:, 0].tolist(), flush=True)
    import time
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200):
        hdl.barrier(channel=0)
    torch.cuda.synchronize()
    print(rank, "barrier us", (time.perf_counter() - t0) / 200 * 1e6, flush=True)
    x = torch.zeros(world * 4096, dtype=torch.float64, device=dev); y = torch.ones(4096, dtype=torch.float64, device=dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200):
        dist.all_gather_into_tensor(x, y)
    torch.cuda.synchronize()
    print(rank, "nccl all_gather 32KB us", (time.perf_counter() - t0) / 200 * 1e6, flush=True)
except Exception as e:
    print(rank, "FAILED", type(e).__name__, str(e)[:300], flush=True)
dist.destroy_process_group()
