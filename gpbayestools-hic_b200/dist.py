"""
Multi-GPU evaluation: walkers shard across the GPUs of one box, one process per GPU
(torch.distributed, NCCL over NVLink/NVSwitch).  Rows are independent (src/mcmc.py:293,
src/emulator.py:584), so the only collective on the data path is the all-gather of the per-walker
log-posteriors (8 bytes per walker); emulator state is replicated on every GPU at load time.

The evaluator is a callable so the partition / gather logic can be exercised on CPU with gloo
(tests/test_dist_cpu.py); on a GPU box it is DeviceChain.log_target_device.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(N, world, rank):
    """Contiguous row block [lo, hi) of rank `rank`: ceil(N/world) rows each, the tail may be
    short or empty."""
    per = -(-N // world) if N > 0 else 0
    lo = min(N, rank * per)
    return lo, min(N, lo + per), per


class PeerGather:
    """All-gather of the per-walker results WITHOUT a collective call: every rank owns a symmetric
    buffer that all other ranks of the box have mapped (torch symmetric memory: cuMem + NVLink peer
    access); the last kernel of the log-posterior path stores each value straight into all of them
    (gpbt_log_posterior_scatter) and one device-side barrier (~7 us) orders the ranks.  Two halves
    alternate between calls so a fast rank never overwrites what a peer may still be reading.
    Replaces an NCCL all-gather whose cost is pure latency (8 bytes per walker)."""

    def __init__(self, rows_per_rank, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.rows = int(rows_per_rank)
        self.buf = symm_mem.empty(2 * self.world * self.rows, dtype=torch.float64, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group.group_name)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.parity = 0
        self.buf.zero_()
        self.hdl.barrier(channel=0)

    def evaluate(self, chain, X_local, oob_value, path=None, mid_event=None):
        """chain: DeviceChain.  Returns the gathered vector [world * n_local] (a view of this rank's
        symmetric buffer, valid until the call after next).  mid_event (a torch.cuda.Event) is recorded
        between the kernels and the device barrier, for callers that time the two apart."""
        n = X_local.shape[0]
        if n > self.rows:
            raise ValueError("PeerGather was sized for %d rows per rank, got %d" % (self.rows, n))
        half = self.parity * self.world * self.rows
        self.parity ^= 1
        peers = [p + 8 * half for p in self.ptrs]
        chain.log_target_scatter(X_local, oob_value, peers, self.rank * n, path=path)
        if mid_event is not None:
            mid_event.record()
        self.hdl.barrier(channel=0)
        return self.buf[half:half + self.world * n]


class ShardedEvaluator:
    def __init__(self, eval_fn, device, group=None):
        """eval_fn(X_local [n_local, p] tensor on `device`) -> lp_local [n_local] float64 tensor."""
        import torch.distributed as dist
        self.eval_fn = eval_fn
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def gather_shards(self, lp_local, per):
        """all-gather equally sized (padded) shards -> [world * per]"""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return lp_local
        if lp_local.shape[0] != per:
            pad = torch.zeros(per, dtype=lp_local.dtype, device=lp_local.device)
            pad[:lp_local.shape[0]] = lp_local
            lp_local = pad
        out = torch.empty(self.world * per, dtype=lp_local.dtype, device=lp_local.device)
        dist.all_gather_into_tensor(out, lp_local.contiguous(), group=self.group)
        return out

    def evaluate_local(self, X_local):
        """Weak-scaling form: every rank already holds its own rows; returns the gathered vector
        [world * n_local] (all ranks must pass the same n_local)."""
        lp = self.eval_fn(X_local)
        return self.gather_shards(lp, X_local.shape[0])

    def evaluate(self, X_host, src=0):
        """Sampler form: rank `src` holds X [N, p] (others may pass None or the same array); X is
        broadcast, each rank evaluates its row block, the result is all-gathered and returned as a
        NumPy array [N] on every rank."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            X_d = torch.from_numpy(np.ascontiguousarray(X_host, dtype=np.float64)).to(self.device, non_blocking=True)
            return self.eval_fn(X_d).cpu().numpy()
        shape = torch.zeros(2, dtype=torch.int64, device=self.device)
        if self.rank == src:
            X_host = np.ascontiguousarray(np.array(X_host, dtype=np.float64, ndmin=2))
            shape[0], shape[1] = X_host.shape
        dist.broadcast(shape, src=src, group=self.group)
        N, p = int(shape[0]), int(shape[1])
        if self.rank == src:
            X_d = torch.from_numpy(X_host).to(self.device)
        else:
            X_d = torch.empty((N, p), dtype=torch.float64, device=self.device)
        dist.broadcast(X_d, src=src, group=self.group)
        lo, hi, per = shard_bounds(N, self.world, self.rank)
        if hi > lo:
            lp_local = self.eval_fn(X_d[lo:hi].contiguous())
        else:
            lp_local = torch.empty(0, dtype=torch.float64, device=self.device)
        return self.gather_shards(lp_local, per)[:N].cpu().numpy()
