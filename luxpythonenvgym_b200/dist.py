"""Multi-GPU plumbing: matches shard independently over ranks; the only collective is the
all-reduce of the episode-statistics vector.  One process per GPU."""
import numpy as np

from . import state as S

STATS_LEN = 16


def shard_range(n_total, rank, world):
    """Contiguous block of match indices [lo, hi) owned by `rank`."""
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi


def stats_from_headers(hdr, size):
    """Host statement of lux_b200_stats (include/lux_b200.h) on a structured header array."""
    hdr = np.asarray(hdr).view(S.header_dtype).reshape(-1)
    done = hdr["done"] != 0
    out = np.zeros(STATS_LEN, dtype=np.int64)
    out[0] = (~done).sum()
    out[1] = done.sum()
    out[2] = (done & (hdr["winner"] == S.WIN_A)).sum()
    out[3] = (done & (hdr["winner"] == S.WIN_B)).sum()
    out[4] = (done & (hdr["winner"] == S.WIN_TIE)).sum()
    out[5] = hdr["turn"].astype(np.int64).sum()
    out[6] = hdr["fuel_generated"].astype(np.int64).sum()
    out[7] = hdr["n_units"].astype(np.int64).sum()
    out[8] = hdr["n_tiles"].astype(np.int64).sum()
    out[9] = hdr["n_cities"].astype(np.int64).sum()
    out[10] = (hdr["status"] != 0).sum()
    live = ~done
    u, k, t = (hdr[f].astype(np.int64) for f in ("n_units", "n_cities", "n_tiles"))
    out[11] = (live * (2 * (8 * size * size + 16 * u + 16 * k + 64) + 4 * (u + t))).sum()
    return out


def all_reduce_stats(stats):
    """Sum the statistics vector over ranks (NCCL on GPU tensors, gloo on CPU tensors)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def all_reduce_max(values):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(values, op=dist.ReduceOp.MAX)
    return values
