"""Probe: several carts stacked on one open cell -- do the kernels build the road as the sequential engine does?"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from tests import test_status_paths as T
from oracle import coracle
from luxpythonenvgym_b200.batched import BatchedGame
lines = ["rp 0 0", "rp 1 0", "r wood 20 20 400",
         "u 1 0 u_1 10 11 0 0 0 0", "u 1 0 u_2 10 11 0 0 0 0", "u 1 1 u_3 10 11 0 0 0 0", "u 0 0 u_4 5 5 0 0 0 0", "u 0 1 u_5 6 6 0 0 0 0",
         "c 0 c_1 400 23", "ct 0 c_1 3 4 0", "c 1 c_2 400 23", "ct 1 c_2 28 28 0"]
ms = T.build(lines, uid=8, cid=2)
a = ms.copy()
g = BatchedGame(4, T.SIZE, caps=T.CAPS)
g.load_states([ms.copy() for _ in range(4)])
uw = np.zeros(T.CAPS["units"], dtype=np.uint32); tb = np.zeros(T.CAPS["tiles"], dtype=np.uint8)
c = 11 * 32 + 10
for t in range(3):
    coracle.step(a, uw, tb)
    g.unit_actions.zero_(); g.tile_actions.zero_(); g.step()
    b = g.export_states()[0]
    print("turn", t, "diff", a.diff(b)[:2], "oracle road_q", a.cells[c]["road_q"], "gpu road_q", b.cells[c]["road_q"], "status", int(b.hdr["status"]), "stats", a.hdr["roads_built_q"], b.hdr["roads_built_q"])
