"""world_size-2 gloo test of the N>1 path's host logic: shard ranges, per-rank stepping of an
independent shard, all-reduce of the statistics vector.  (On GPUs the same code runs with NCCL;
here the per-rank stepping uses the oracle because the container has no GPU.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from luxpythonenvgym_b200 import dist as ldist
from luxpythonenvgym_b200 import mapgen

N_TOTAL, SIZE, TURNS, SEED = 24, 12, 45, 99


def _run_shard(lo, hi):
    from oracle import coracle
    hdrs = []
    for m in range(lo, hi):
        ms = mapgen.generate(3000 + m, size=SIZE)
        for _ in range(TURNS):
            uw, tb = coracle.gen_actions(ms, SEED, m)
            coracle.step(ms, uw, tb)
        hdrs.append(ms.hdr.copy())
    return np.array(hdrs)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = ldist.shard_range(N_TOTAL, rank, world)
    stats = torch.from_numpy(ldist.stats_from_headers(_run_shard(lo, hi), SIZE))
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    ldist.all_reduce_stats(stats)
    ldist.all_reduce_max(ms)
    if rank == 0:
        out.put((stats.numpy().tolist(), ms.item()))
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    for n, world in ((24, 2), (1_000_000, 8), (7, 3), (5, 8)):
        spans = [ldist.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_two_ranks_reduce_to_single_process_result():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    stats, ms = out.get()
    want = ldist.stats_from_headers(_run_shard(0, N_TOTAL), SIZE)
    assert stats == want.tolist()
    assert ms == 11.0
    assert want[0] + want[1] == N_TOTAL
