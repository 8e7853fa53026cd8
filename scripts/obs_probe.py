"""Observation kernel time against the batch capacities (its shared-memory footprint is laid out for the capacities):
python scripts/obs_probe.py"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from luxpythonenvgym_b200 import mapgen
from luxpythonenvgym_b200.batched import BatchedGame
N, SIZE = 16384, 32
for caps in (None, dict(units=128, cities=128, tiles=192, res=192), dict(units=64, cities=64, tiles=96, res=192)):
    g = BatchedGame(N, SIZE, caps); pool = BatchedGame(256, SIZE, caps)
    pool.load_arrays(mapgen.generate_batch(np.arange(424242, 424242 + 256), SIZE, caps=caps))
    g.load_arrays(mapgen.generate_batch(np.arange(1000000, 1000000 + N), SIZE, caps=caps))
    resets = torch.zeros(1, dtype=torch.int32, device="cuda")
    g.gen_actions(7, 0)
    for t in range(150):
        g.step(); g.reset_gen(pool, t, 7, 0, reset_count=resets)
    sel = torch.zeros(N, dtype=torch.uint8, device="cuda")
    for _ in range(5): g.observe(sel, 32)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): g.observe(sel, 32)
    e1.record(); torch.cuda.synchronize()
    st = g.stats().cpu().numpy()
    print("caps %s: observe (padded, e_cap 32) %.4f ms; units/match %.1f status_nonzero %d" % (g.caps, e0.elapsed_time(e1) / 50, st[7] / N, st[10]))
