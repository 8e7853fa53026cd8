"""Timeline of one step's launches on the bench workload (globaltimer stamps, option "trace"):
  python scripts/trace_step.py "only_class=0" "only_class=2" "slice=5120" ...      (each argument: options for one line)"""
import ctypes, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from luxpythonenvgym_b200 import _lib, mapgen
from luxpythonenvgym_b200.batched import BatchedGame
lib = _lib.load()
N, SIZE = 16384, 32
game = BatchedGame(N, SIZE); pool = BatchedGame(1024, SIZE)
pool.load_arrays(mapgen.generate_batch(np.arange(424242, 424242 + 1024), SIZE))
game.load_arrays(mapgen.generate_batch(np.arange(1000000, 1000000 + N), SIZE))
resets = torch.zeros(1, dtype=torch.int32, device="cuda")
game.gen_actions(20211019, 0)
for t in range(400):
    game.step(); game.reset_gen(pool, t, 20211019, 0, reset_count=resets)
torch.cuda.synchronize()
lib.lux_b200_set_option(b"trace", 1)
out = (ctypes.c_ulonglong * 40)()
t = 400
for spec in sys.argv[1:] or ["only_class=0"]:
    for kv in spec.split(","):
        k, v = kv.split("="); lib.lux_b200_set_option(k.encode(), int(v))
    rows = []
    for rep in range(12):
        game.step()
        game.reset_gen(pool, t, 20211019, 0, reset_count=resets); t += 1
        lib.lux_b200_debug_trace(out)
        v = list(out)
        hist = v[8:40]
        t0 = v[6]
        rows.append([(v[2 * i] - t0) / 1e3 if v[2 * i + 1] else float("nan") for i in (3, 0, 1, 2)] + [(v[2 * i + 1] - t0) / 1e3 if v[2 * i + 1] else float("nan") for i in (3, 0, 1, 2)])   # columns: classify, big, harness (reset + action stream), small
    r = np.nanmedian(np.array(rows[2:]), axis=0)
    print("   small-class warps starting per 8 us bin:", hist[:16], " big-class warps ending per bin:", hist[16:])
    print("%-40s classify %.1f-%.1f  big class %.1f-%.1f  small class %.1f-%.1f  reset + action stream %.1f-%.1f us" % (spec, r[0], r[4], r[1], r[5], r[3], r[7], r[2], r[6]))
