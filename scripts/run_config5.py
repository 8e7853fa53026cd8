"""BASELINE.json config 5, literally: 1M independent seeded 32x32 matches played to the end, sharded
contiguously over the ranks of one box, episode statistics reduced with one NCCL all-reduce.

  python scripts/run_config5.py [--matches 1048576] [--chunk 16384]                      (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29531 scripts/run_config5.py                                         (8 GPUs)

Each rank owns seeds [lo, hi) (luxpythonenvgym_b200.dist.shard_range), generates its maps with the native
host generator while the GPU plays the previous chunk, keeps `chunk` matches resident, and steps a chunk
until every match in it is over (canonical seeded action stream).  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from luxpythonenvgym_b200 import dist as ldist  # noqa: E402
from luxpythonenvgym_b200 import mapgen  # noqa: E402
from luxpythonenvgym_b200.batched import BatchedGame  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--matches", type=int, default=1 << 20)
    ap.add_argument("--chunk", type=int, default=16384)
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--seed", type=int, default=2021)
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    lo, hi = ldist.shard_range(args.matches, rank, world)
    threads = max(1, (os.cpu_count() or 8) // world)
    game = BatchedGame(args.chunk, args.size, device=dev)
    totals = torch.zeros(ldist.STATS_LEN, dtype=torch.int64, device=dev)
    counters = torch.zeros(2, dtype=torch.int64, device=dev)
    t_map = [0.0]

    def gen(first):
        t0 = time.perf_counter()
        k = min(args.chunk, hi - first)
        arr = mapgen.generate_batch(np.arange(first, first + k), args.size, threads=threads)
        t_map[0] += time.perf_counter() - t0
        return arr

    pending = {"arr": gen(lo)}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_ms, steps = 0.0, 0
    wall0 = time.perf_counter()
    first = lo
    while first < hi:
        arr = pending["arr"]
        k = arr["hdr"].shape[0]
        game.hdr[:, 24] = 1                       # slots beyond k stay "done"
        game.load_arrays(arr, 0)
        nxt = first + k
        worker = None
        if nxt < hi:                              # overlap host map generation with GPU play
            worker = threading.Thread(target=lambda: pending.__setitem__("arr", gen(nxt)))
            worker.start()
        e0.record()
        for turn in range(360):
            game.gen_actions(args.seed, first, counters=counters)
            game.step()
            steps += 1
            if turn >= 30 and turn % 30 == 0 and not bool((game.hdr[:k, 24] == 0).any()):
                break
        e1.record()
        torch.cuda.synchronize()
        dev_ms += e0.elapsed_time(e1)
        st = game.stats()
        live_pad = args.chunk - k                 # padding slots are counted as finished by lux_b200_stats
        st = st.clone()
        st[1] -= live_pad
        totals += st
        if worker:
            worker.join()
        first = nxt
    wall = time.perf_counter() - wall0
    times = torch.tensor([dev_ms, wall * 1e3, t_map[0] * 1e3], dtype=torch.float64, device=dev)
    ldist.all_reduce_stats(totals)
    ldist.all_reduce_stats(counters)
    ldist.all_reduce_max(times)
    if rank == 0:
        t = totals.cpu().tolist()
        live_turns = int(counters[0].item())
        print(json.dumps({
            "config": "1M independent %dx%d matches, %d GPUs, %d resident per GPU" % (args.size, args.size, world, args.chunk),
            "matches": args.matches, "finished": t[1], "wins_A": t[2], "wins_B": t[3], "ties": t[4],
            "match_turns": live_turns, "mean_turns_per_match": live_turns / max(args.matches, 1),
            "sum_fuel_generated": t[6], "status_nonzero": t[10],
            "device_seconds_max_rank": times[0].item() / 1e3, "wall_seconds_max_rank": times[1].item() / 1e3,
            "host_mapgen_seconds_max_rank": times[2].item() / 1e3,
            "match_turns_per_s_device": live_turns / (times[0].item() / 1e3),
            "match_turns_per_s_wall": live_turns / (times[1].item() / 1e3)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
