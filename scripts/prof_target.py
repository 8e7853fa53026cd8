"""Profiling target: the bench workload (32x32, 16384 matches, 400 burn-in turns) followed by a few steps.
  python scripts/prof_target.py [steps]      -- run under ncu with  -k regex:lux_step_kernel -s 1209 -c 3
(400 burn-in + 3 warm-up steps = 1209 lux_step_kernel launches: big class, small class, oversize launch per step)."""
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from luxpythonenvgym_b200 import mapgen
from luxpythonenvgym_b200.batched import BatchedGame

import os
from luxpythonenvgym_b200 import _lib
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N, SIZE = 16384, int(os.environ.get("LUX_SIZE", "32"))
for opt in os.environ.get("LUX_OPTS", "").split(","):       # e.g. LUX_OPTS=group=16,carveout=100
    if "=" in opt:
        k, v = opt.split("=")
        assert _lib.load().lux_b200_set_option(k.encode(), int(v)) == 0, opt
game = BatchedGame(N, SIZE)
pool = BatchedGame(1024, SIZE)
pool.load_arrays(mapgen.generate_batch(np.arange(424242, 424242 + 1024), SIZE))
game.load_arrays(mapgen.generate_batch(np.arange(1000000, 1000000 + N), SIZE))
resets = torch.zeros(1, dtype=torch.int32, device="cuda")
game.gen_actions(20211019, 0)
for t in range(400 + 3 + steps):
    game.step()
    game.reset_gen(pool, t, 20211019, 0, reset_count=resets)
torch.cuda.synchronize()
print("ok", int(game.stats()[0]))
