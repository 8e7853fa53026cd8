"""Tuning probe (not the contract bench): the bench workload (32x32, 16384 matches, 400 burn-in turns) replayed from
one snapshot under several lux_b200_set_option settings.
  python scripts/sweep.py "group=32" "group=16,carveout=100" ...      (each argument = one configuration)
Environment: LUX_SWEEP_N (matches, default 16384), LUX_SWEEP_STEPS (default 60), LUX_SWEEP_DENSE=1 (dense replay pool)."""
import os
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from luxpythonenvgym_b200 import _lib, mapgen
from luxpythonenvgym_b200.batched import BatchedGame

N = int(os.environ.get("LUX_SWEEP_N", "16384"))
STEPS = int(os.environ.get("LUX_SWEEP_STEPS", "60"))
SIZE = int(os.environ.get("LUX_SWEEP_SIZE", "32"))
DEFAULTS = {"only_class": 0, "big_share": 3, "tma": 1, "warps_per_cta": 0, "two_class": -1, "list_grid": 0, "overlap": 1, "group": 0, "carveout": -1, "slice": 0, "fork": 1, "mid_slice": 0, "over_own": -1, "cta_sync": -1, "sync_mask": 84, "order": 1, "n_ord": 0, "ord_step": 10}

lib = _lib.load()
pool_arrays = mapgen.generate_batch(np.arange(424242, 424242 + 1024), SIZE)
game = BatchedGame(N, SIZE)
pool = BatchedGame(1024, SIZE)
pool.load_arrays(pool_arrays)
if os.environ.get("LUX_SWEEP_DENSE"):
    from tests import golden_util as G
    z = G.npz("dense")
    states = [G.get_state(z, name + "/init") for name in G.dense_names()]
    from luxpythonenvgym_b200 import state as S
    caps = S.default_caps(SIZE)

    def recap(ms):
        o = S.MatchState(ms.size, caps)
        o.hdr, o.cells = ms.hdr.copy(), ms.cells.copy()
        for name in ("units", "cities", "tiles", "res"):
            dst, src = getattr(o, name), getattr(ms, name)
            k = min(len(dst), len(src))
            dst[:k] = src[:k]
        return o

    states = [recap(s) for s in states if s.size == SIZE]
    pool = BatchedGame(len(states), SIZE)
    pool.load_states(states)
    game.load_states([states[i % len(states)] for i in range(N)])
    burn = 20
else:
    game.load_arrays(mapgen.generate_batch(np.arange(1000000, 1000000 + N), SIZE))
    burn = 400
resets = torch.zeros(1, dtype=torch.int32, device="cuda")
game.gen_actions(20211019, 0)
for t in range(burn):
    game.step()
    game.reset_gen(pool, t, 20211019, 0, reset_count=resets)
torch.cuda.synchronize()
snap = {k: t.clone() for k, t in game.sections().items()}
snap_ua, snap_ta = game.unit_actions.clone(), game.tile_actions.clone()
st = game.stats().cpu().numpy()
print("workload: %d matches %dx%d, live %d, units/match %.1f tiles/match %.1f" % (N, SIZE, SIZE, st[0], st[7] / N, st[8] / N))

ref_hdr = None
for spec in sys.argv[1:] or ["group=32"]:
    opts = dict(DEFAULTS)
    for kv in spec.split(","):
        if "=" in kv:
            k, v = kv.split("=")
            opts[k] = int(v)
    for k, v in opts.items():
        lib.lux_b200_set_option(k.encode(), v)
    for k, t in game.sections().items():
        t.copy_(snap[k])
    game.unit_actions.copy_(snap_ua)
    game.tile_actions.copy_(snap_ta)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for w in range(5):
        game.step()
        game.reset_gen(pool, burn + w, 20211019, 0, reset_count=resets)
    torch.cuda.synchronize()
    t0.record()
    for i in range(STEPS):
        ev[i][0].record()
        game.step()
        ev[i][1].record()
        game.reset_gen(pool, burn + 5 + i, 20211019, 0, reset_count=resets)
    t1.record()
    torch.cuda.synchronize()
    step_ms = sorted(a.elapsed_time(b) for a, b in ev)
    hdr = game.hdr.clone()
    same = "" if ref_hdr is None else (" same-result" if torch.equal(hdr, ref_hdr) else " DIFFERENT-RESULT")
    if ref_hdr is None:
        ref_hdr = hdr
    print("%-44s step median %.4f ms (min %.4f)  whole %.4f ms%s" % (spec, step_ms[len(step_ms) // 2], step_ms[0], t0.elapsed_time(t1) / STEPS, same))
