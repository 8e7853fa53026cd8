"""Rough timing probe (not the contract bench): python scripts/quick_bench.py SIZE N STEPS"""
import sys, time
sys.path.insert(0, ".")
import torch
from luxpythonenvgym_b200 import mapgen
from luxpythonenvgym_b200.batched import BatchedGame

size, n, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
t0 = time.time()
maps = [mapgen.generate(7000 + i, size=size) for i in range(64)]
print("mapgen 64 maps %.2fs" % (time.time() - t0))
g = BatchedGame(n, size)
g.load_states([maps[i % 64] for i in range(n)])
pool = BatchedGame(64, size)
pool.load_states(maps)
count = torch.zeros(1, dtype=torch.int32, device="cuda")
for warm in range(40):
    g.gen_actions(1, 0); g.step(); g.reset_done(pool, warm, count)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tg = ts = tr = 0.0
live = 0
for i in range(steps):
    e[0].record(); g.gen_actions(1, 0); e[1].record(); g.step(); e[2].record(); g.reset_done(pool, 1000 + i, count); e[3].record()
    torch.cuda.synchronize()
    tg += e[0].elapsed_time(e[1]); ts += e[1].elapsed_time(e[2]); tr += e[2].elapsed_time(e[3])
st = g.stats().cpu().numpy()
print("size %d n %d: gen %.3f ms, step %.3f ms, reset %.3f ms per turn; step-only %.3e match-turns/s; all %.3e" % (
    size, n, tg / steps, ts / steps, tr / steps, n / (ts / steps) * 1e3, n / ((tg + ts + tr) / steps) * 1e3))
print("stats live %d done %d sum_units %d sum_tiles %d bad %d bytes/turn %d resets %d" % (st[0], st[1], st[7], st[8], st[10], st[11], int(count.item())))
