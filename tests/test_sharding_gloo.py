"""Multi-rank host logic on CPU: world_size-2 gloo run of the slice-sharding used by bench.py --gpus N.
There is no data-path collective (slices are independent); what is exercised is the bookkeeping around it:
slice bounds, reproducible per-slice input generation, and the max-over-ranks timing reduction."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 1_000_003


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fabe_b200
    from oracle.loader import Oracle

    b, e = fabe_b200.shard_bounds(N, world, rank)
    o = Oracle()
    x = o.fill_uniform(e - b, 0xFABE1303, -1e4, 1e4, first_index=b)  # the slice of the global array this rank owns
    s, c, _ = fabe_b200.host_core(x)  # host instantiation of the core (no GPU here)
    # every rank reports (begin, end, checksum, elapsed); elapsed is reduced with MAX exactly as bench.py does
    info = torch.tensor([b, e], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, info)
    t = torch.tensor([0.5 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    chk = torch.tensor([float(np.abs(s).sum() + np.abs(c).sum())], dtype=torch.float64)
    dist.all_reduce(chk, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(([g.tolist() for g in gathered], t.item(), chk.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    from fabe_b200.build import build_library

    build_library()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    bounds, tmax, chk = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert bounds[0][0] == 0 and bounds[0][1] == bounds[1][0] and bounds[1][1] == N
    assert bounds[0][1] % 4 == 0
    assert tmax == 1.5  # max over ranks
    # the sharded run reproduces the single-process result
    sys.path.insert(0, ROOT)
    import fabe_b200
    from oracle.loader import Oracle

    x = Oracle().fill_uniform(N, 0xFABE1303, -1e4, 1e4)
    s, c, _ = fabe_b200.host_core(x)
    assert abs(chk - float(np.abs(s).sum() + np.abs(c).sum())) < 1e-6
