"""Small deterministic case for compute-sanitizer (memcheck / racecheck): 32 matches of 16x16 and 8 dense
32x32 matches, 25 turns with the canonical action stream, pre-validation on, observations + rewards."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from luxpythonenvgym_b200 import mapgen, state as S  # noqa: E402
from luxpythonenvgym_b200.batched import BatchedGame  # noqa: E402


def main():
    g = BatchedGame(32, 16)
    g.load_arrays(mapgen.generate_batch(np.arange(100, 132), 16))
    z = np.load(os.path.join(ROOT, "tests", "golden", "dense.npz"))
    names = [str(n) for n in z["names"] if int(z[str(n) + "/size"]) == 32][:8]
    caps = S.default_caps(32)
    dense = []
    for nm in names:
        ms = S.MatchState.from_arrays(32, {k: z["%s/init/%s" % (nm, k)] for k in S.MatchState.SECTIONS})
        o = S.MatchState(32, caps)
        o.hdr, o.cells = ms.hdr.copy(), ms.cells.copy()
        for name in ("units", "cities", "tiles", "res"):
            k = min(len(getattr(o, name)), len(getattr(ms, name)))
            getattr(o, name)[:k] = getattr(ms, name)[:k]
        dense.append(o)
    d = BatchedGame(len(dense), 32)
    d.load_states(dense)
    for game in (g, d):
        sel = torch.zeros(game.n, dtype=torch.uint8, device="cuda")
        rstate = torch.zeros((game.n, 4), dtype=torch.int32, device="cuda")
        for turn in range(25):
            game.gen_actions(7, 0)
            game.step(prevalidate=(turn % 2 == 0))
            game.observe(sel)
            game.reward(sel, rstate)
        torch.cuda.synchronize()
        print("ok", game.size, game.stats().cpu().tolist()[:11])


if __name__ == "__main__":
    main()
