"""Rough timing probe (not the contract bench): python scripts/quick_bench.py SIZE N STEPS [caps u,c,t,r] [wpc]"""
import sys, time
sys.path.insert(0, ".")
import torch
from luxpythonenvgym_b200 import mapgen, _lib
from luxpythonenvgym_b200.batched import BatchedGame

size, n, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
caps = None
if len(sys.argv) > 4 and sys.argv[4] != "-":
    u, c, t, r = [int(x) for x in sys.argv[4].split(",")]
    caps = dict(units=u, cities=c, tiles=t, res=r)
wpcs = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0]
maps = [mapgen.generate(7000 + i, size=size, caps=caps) for i in range(64)]
g = BatchedGame(n, size, caps)
g.load_states([maps[i % 64] for i in range(n)])
pool = BatchedGame(64, size, caps)
pool.load_states(maps)
count = torch.zeros(1, dtype=torch.int32, device="cuda")
lib = _lib.load()
for warm in range(400):
    g.gen_actions(1, 0); g.step(); g.reset_done(pool, warm, count)
torch.cuda.synchronize()
print("smem/match", lib.lux_b200_step_smem_bytes(size, g.caps["units"], g.caps["cities"], g.caps["tiles"], g.caps["res"]), g.caps)
import os
for opt in os.environ.get("LUX_OPTS", "").split(","):
    if "=" in opt:
        k, v = opt.split("=")
        print("option", k, v, lib.lux_b200_set_option(k.encode(), int(v)))
for wpc in wpcs:
    lib.lux_b200_set_option(b"warps_per_cta", wpc)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tg = ts = tr = 0.0
    for i in range(steps):
        e[0].record(); g.gen_actions(1, 0); e[1].record(); g.step(); e[2].record(); g.reset_done(pool, 1000 + i, count); e[3].record()
        torch.cuda.synchronize()
        tg += e[0].elapsed_time(e[1]); ts += e[1].elapsed_time(e[2]); tr += e[2].elapsed_time(e[3])
    st = g.stats().cpu().numpy()
    print("wpc %d size %d n %d: gen %.3f ms, step %.3f ms, reset %.3f ms; step-only %.3e mt/s; units/match %.1f status_bad %d" % (
        wpc, size, n, tg / steps, ts / steps, tr / steps, n / (ts / steps) * 1e3, st[7] / n, st[10]))
