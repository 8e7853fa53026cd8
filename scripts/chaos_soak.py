"""Adversarial soak 2: random hand-built states around NIGHTFALL -- multi-tile cities of both teams with little fuel and
many units standing on their tiles (so that cities fall under several units and stacks form by themselves), carts and
workers with cargo, resources and roads around -- played with the canonical random action stream, GPU vs the C oracle
on every turn (matches the engine flags LUX_ST_BAD_STATE are left alone).
  python scripts/chaos_soak.py [n_states] [turns] [seed]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from tests import test_status_paths as T
from oracle import coracle
from luxpythonenvgym_b200.batched import BatchedGame

n_states = int(sys.argv[1]) if len(sys.argv) > 1 else 256
turns = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
# launcher options for the run, e.g. LUX_OPTS=cta_sync=8,sync_mask=511,slice=3200,two_class=1,mid_slice=6144
import os
if os.environ.get("LUX_OPTS"):
    from luxpythonenvgym_b200 import _lib as _opt_lib
    for _kv in os.environ["LUX_OPTS"].split(","):
        _k, _v = _kv.split("=")
        assert _opt_lib.load().lux_b200_set_option(_k.encode(), int(_v)) == 0, _kv
states = []
for k in range(n_states):
    used = set()
    lines = ["rp 0 %d" % rng.integers(0, 260), "rp 1 %d" % rng.integers(0, 260)]
    def free(near=None):
        for _ in range(200):
            if near is None:
                p = (int(rng.integers(1, 31)), int(rng.integers(1, 31)))
            else:
                p = (near[0] + int(rng.integers(-2, 3)), near[1] + int(rng.integers(-2, 3)))
            if p not in used and 0 <= p[0] < 32 and 0 <= p[1] < 32:
                used.add(p); return p
        return None
    uid, cid = 1, 1
    city_lines, tile_cells = [], []
    for team in (0, 1):
        for _ in range(int(rng.integers(1, 4))):
            start = free()
            cells = [start]
            for _ in range(int(rng.integers(0, 6))):
                b = cells[int(rng.integers(0, len(cells)))]
                d = [(1, 0), (-1, 0), (0, 1), (0, -1)][int(rng.integers(0, 4))]
                p = (b[0] + d[0], b[1] + d[1])
                if p not in used and 0 <= p[0] < 32 and 0 <= p[1] < 32:
                    used.add(p); cells.append(p)
            fuel = int(rng.choice([0, 5, 20, 60, 300, 2000]))
            city_lines.append("c %d c_%d %d %d" % (team, cid, fuel, 23 * len(cells)))
            for (x, y) in cells:
                city_lines.append("ct %d c_%d %d %d %d" % (team, cid, x, y, int(rng.integers(0, 3))))
                tile_cells.append((team, x, y))
            cid += 1
    for _ in range(int(rng.integers(8, 30))):
        p = free(near=tile_cells[int(rng.integers(0, len(tile_cells)))][1:] if rng.random() < 0.6 else None)
        if p is None: continue
        kind = ["wood", "coal", "uranium"][int(rng.integers(0, 3))]
        lines.append("r %s %d %d %d" % (kind, p[0], p[1], int(rng.integers(1, 500))))
    unit_lines = []
    for (team, x, y) in tile_cells:
        for _ in range(int(rng.integers(0, 4))):
            typ = int(rng.random() < 0.35)
            cap = 2000 if typ else 100
            # cargo enough to survive a night or not
            w = int(rng.choice([0, 0, 3, 8, 40, cap // 2])); c = int(rng.choice([0, 0, 1, 5])); u = int(rng.choice([0, 0, 0, 2]))
            unit_lines.append("u %d %d u_%d %d %d %d %d %d %d" % (typ, team, uid, x, y, int(rng.integers(0, 3)), w, c, u)); uid += 1
    for _ in range(int(rng.integers(0, 8))):
        p = free()
        if p is None: continue
        typ = int(rng.random() < 0.3)
        unit_lines.append("u %d %d u_%d %d %d 0 %d 0 0" % (typ, int(rng.integers(0, 2)), uid, p[0], p[1], int(rng.integers(0, 100)))); uid += 1
    road_lines = []
    for _ in range(int(rng.integers(0, 6))):
        p = free()
        if p is not None:
            road_lines.append("ccd %d %d %s" % (p[0], p[1], ["0.75", "3", "4.5", "5.25", "6"][int(rng.integers(0, 5))]))
    turn = int(rng.integers(0, 8)) * 40 + int(rng.integers(24, 40))
    try:
        states.append(T.build(lines + unit_lines + city_lines + road_lines, turn=min(turn, 350), uid=uid + 1, cid=cid + 1))
    except AssertionError:
        pass
n_states = len(states)
traj = [s.copy() for s in states]
game = BatchedGame(n_states, T.SIZE, caps=T.CAPS)
game.load_states(states)
bad, flagged, stacks_seen = 0, set(), 0
for t in range(turns):
    uws, tbs = [], []
    prev = [ms.copy() for ms in traj]
    for m, ms in enumerate(traj):
        uw, tb = coracle.gen_actions(ms, 777, m)
        uws.append(uw); tbs.append(tb)
        coracle.step(ms, uw, tb)
    game.unit_actions.copy_(torch.from_numpy(np.stack(uws).view(np.int32)))
    game.tile_actions.copy_(torch.from_numpy(np.stack(tbs)))
    game.step()
    got = game.export_states()
    for m in range(n_states):
        if int(got[m].hdr["status"]) & 8:
            flagged.add(m); continue
        if m in flagged: continue
        d = traj[m].diff(got[m])
        if d or int(got[m].hdr["status"]) != int(traj[m].hdr["status"]):
            bad += 1
            if bad <= 6:
                print("turn %d (game turn %d) match %d: %s (status gpu %d oracle %d)" % (t, int(prev[m].hdr["turn"]), m, d[:3], int(got[m].hdr["status"]), int(traj[m].hdr["status"])))
    # how many non-city cells hold several units (stacks that formed by themselves)
    for m in range(0, n_states, 8):
        ms = traj[m]; nu = int(ms.hdr["n_units"])
        if nu < 2: continue
        pos = ms.units["y"][:nu].astype(np.int32) * 32 + ms.units["x"][:nu]
        open_cell = (ms.cells["flags"][pos] & 0x0C) == 0
        vals, cnt = np.unique(pos[open_cell], return_counts=True)
        stacks_seen += int((cnt > 1).sum())
    if bad: break
live = sum(1 for ms in traj if not int(ms.hdr["done"]))
print("chaos soak: %d states x %d turns, mismatching match-turns: %d, flagged LUX_ST_BAD_STATE: %d, stack cells seen (1/8 sample): %d, still live: %d" % (n_states, t + 1, bad, len(flagged), stacks_seen, live))
