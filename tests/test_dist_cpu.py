"""world_size-2 gloo tests (CPU) of the walker sharding / all-gather logic in gpbt_b200.dist."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gpbt_b200  # noqa: F401
        from gpbt_b200.dist import ShardedEvaluator, shard_bounds
        calls = []

        def fake_eval(X):  # stands in for DeviceChain.log_target_device
            calls.append(X.shape[0])
            return X.sum(dim=1) * 2.0

        ev = ShardedEvaluator(fake_eval, torch.device("cpu"))
        rng = np.random.default_rng(5)
        X = rng.normal(size=(N, 3))
        out = ev.evaluate(X if rank == 0 else None, src=0)
        lo, hi, per = shard_bounds(N, world, rank)
        res = dict(ok=bool(np.allclose(out, X.sum(1) * 2.0)) and out.shape == (N,),
                   rows=list(calls), bounds=(lo, hi, per))
        # weak-scaling form
        Xl = torch.full((4, 3), float(rank + 1), dtype=torch.float64)
        g = ev.evaluate_local(Xl)
        res["weak"] = g.tolist()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N", [10, 7, 1, 0])
def test_sharded_evaluate_gloo(N):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert got[r]["ok"], got[r]
        assert got[r]["weak"] == [6.0] * 4 + [12.0] * 4
    per = -(-N // world) if N else 0
    assert got[0]["bounds"] == (0, min(N, per), per)
    assert sum(sum(got[r]["rows"]) for r in range(world)) == N


def test_shard_bounds_cover():
    import gpbt_b200  # noqa: F401
    from gpbt_b200.dist import shard_bounds
    for N in (0, 1, 5, 4096, 4097):
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                lo, hi, per = shard_bounds(N, world, r)
                assert 0 <= hi - lo <= per
                rows += list(range(lo, hi))
            assert rows == list(range(N))
