"""Host-side breakdown of the end-to-end list path (bench.py's e2e loop): time spent in submit / reset / wait per step."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from luxpythonenvgym_b200 import mapgen
from luxpythonenvgym_b200.batched import BatchedGame
N, SIZE, E = 16384, 32, 300
game = BatchedGame(N, SIZE); pool = BatchedGame(1024, SIZE)
pool.load_arrays(mapgen.generate_batch(np.arange(424242, 424242 + 1024), SIZE))
game.load_arrays(mapgen.generate_batch(np.arange(1000000, 1000000 + N), SIZE))
resets = torch.zeros(1, dtype=torch.int32, device="cuda")
game.gen_actions(20211019, 0)
for t in range(400):
    game.step(); game.reset_gen(pool, t, 20211019, 0, reset_count=resets)
snap = {k: v.clone() for k, v in game.sections().items()}
ents, ks = [], []
for k in range(E):
    game.unit_actions.zero_(); game.tile_actions.zero_(); game.gen_actions(20211019, 0)
    e = game.pack_entries(); ents.append(e.cpu().pin_memory()); ks.append(int(e.shape[0]))
    game.step(); game.reset_done(pool, 400 + k, resets)
for k, v in game.sections().items(): v.copy_(snap[k])
game.unit_actions.zero_(); game.tile_actions.zero_()
done_h = torch.empty(N, dtype=torch.uint8).pin_memory(); summ_h = torch.zeros(4, dtype=torch.int32).pin_memory()
torch.cuda.synchronize()
ts = np.zeros((E, 4))
for k in range(E):
    t0 = time.perf_counter(); game.step_host_sparse_submit(ents[k], ks[k])
    t1 = time.perf_counter(); game.reset_done(pool, 400 + k, resets)
    t2 = time.perf_counter(); game.step_host_wait(done_h, summ_h)
    t3 = time.perf_counter(); ts[k] = (t1 - t0, t2 - t1, t3 - t2, t3 - t0)
torch.cuda.synchronize()
m = np.median(ts[20:], axis=0) * 1e6
print("per step (median, us): submit %.1f  reset_done %.1f  wait %.1f  total %.1f  -> %.3e match-turns/s" % (m[0], m[1], m[2], m[3], N / (m[3] * 1e-6)))
