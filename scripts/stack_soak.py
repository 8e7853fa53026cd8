"""Adversarial soak: random hand-built states with STACKS (several units of both teams and kinds on open cells next to
resources and on roads) played for some turns with the canonical random action stream, GPU vs the C oracle on every
turn.  python scripts/stack_soak.py [n_states] [turns] [seed]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from tests import test_status_paths as T
from oracle import coracle
from luxpythonenvgym_b200.batched import BatchedGame

n_states = int(sys.argv[1]) if len(sys.argv) > 1 else 256
turns = int(sys.argv[2]) if len(sys.argv) > 2 else 30
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
    def free():
        while True:
            p = (int(rng.integers(2, 30)), int(rng.integers(2, 30)))
            if p not in used:
                used.add(p); return p
    for _ in range(int(rng.integers(6, 25))):
        x, y = free()
        kind = ["wood", "coal", "uranium"][int(rng.integers(0, 3))]
        lines.append("r %s %d %d %d" % (kind, x, y, int(rng.integers(1, 500))))
    uid = 1
    stacks = []
    for _ in range(int(rng.integers(1, 5))):
        x, y = free(); stacks.append((x, y))
        for _ in range(int(rng.integers(2, 7))):
            typ, team = int(rng.integers(0, 2)), int(rng.integers(0, 2))
            cap = 2000 if typ else 100
            w = int(rng.integers(0, cap // 2)); c = int(rng.integers(0, max(1, (cap - w) // 3))); u = int(rng.integers(0, max(1, (cap - w - c) // 4)))
            cd = int(rng.integers(0, 3))
            lines.append("u %d %d u_%d %d %d %d %d %d %d" % (typ, team, uid, x, y, cd, w, c, u)); uid += 1
        if rng.random() < 0.6:
            lines.append("ccd %d %d %s" % (x, y, ["0.75", "3", "4.5", "5.25", "5.5", "6"][int(rng.integers(0, 6))]))
    for _ in range(int(rng.integers(0, 6))):
        x, y = free()
        lines.append("u %d %d u_%d %d %d 0 %d 0 0" % (int(rng.integers(0, 2)), int(rng.integers(0, 2)), uid, x, y, int(rng.integers(0, 100)))); uid += 1
    cx, cy = free(); dx, dy = free()
    lines += ["c 0 c_1 %d 23" % rng.integers(0, 600), "ct 0 c_1 %d %d 0" % (cx, cy), "c 1 c_2 %d 23" % rng.integers(0, 600), "ct 1 c_2 %d %d 0" % (dx, dy)]
    states.append(T.build(lines, turn=int(rng.integers(0, 350)), uid=uid + 1, cid=2))
traj = [s.copy() for s in states]
prev = None
flagged = set()
game = BatchedGame(n_states, T.SIZE, caps=T.CAPS)
game.load_states(states)
bad = 0
for t in range(turns):
    uws, tbs = [], []
    prev = [ms.copy() for ms in traj]
    for m, ms in enumerate(traj):
        uw, tb = coracle.gen_actions(ms, 4242, m)
        uws.append(uw); tbs.append(tb)
        coracle.step(ms, uw, tb)
    game.unit_actions.copy_(torch.from_numpy(np.stack(uws).view(np.int32)))
    game.tile_actions.copy_(torch.from_numpy(np.stack(tbs)))
    game.step()
    got = game.export_states()
    for m in range(n_states):
        if int(got[m].hdr["status"]) & 8:          # LUX_ST_BAD_STATE: the engine says the match left its domain
            if m not in flagged:
                flagged.add(m)
            continue
        d = traj[m].diff(got[m])
        if d or int(got[m].hdr["status"]) != int(traj[m].hdr["status"]):
            bad += 1
            if bad <= 8:
                print("turn %d match %d: %s (status gpu %d oracle %d)" % (t, m, d[:2], int(got[m].hdr["status"]), int(traj[m].hdr["status"])))
                bc = np.nonzero(traj[m].cells["road_q"] != got[m].cells["road_q"])[0]
                for c in bc[:2]:
                    x, y = int(c) % T.SIZE, int(c) // T.SIZE
                    print("   cell (%d,%d): road before %d, oracle %d, gpu %d" % (x, y, prev[m].cells[c]["road_q"], traj[m].cells[c]["road_q"], got[m].cells[c]["road_q"]))
                    nu = int(prev[m].hdr["n_units"])
                    for i in range(nu):
                        U = prev[m].units[i]
                        if abs(int(U["x"]) - x) + abs(int(U["y"]) - y) <= 1:
                            w = int(uws[m][i])
                            print("     unit %d id %d at (%d,%d) type/team %d cd %d action kind %d dir %d" % (i, U["id"], U["x"], U["y"], U["team_type"], U["cd_q"], w & 7, (w >> 3) & 7))
    if bad:
        break
print("soak: %d states x %d turns, mismatching match-turns: %d, matches flagged LUX_ST_BAD_STATE (left alone): %d" % (n_states, t + 1, bad, len(flagged)))
