"""Capacity overflows (LUX_ST_UNIT / CITY / TILE_OVERFLOW) and the stacked-units state, on hand-built matches.

The packed state has capacities the reference's Python objects do not (u_cap <= 252 is an ABI limit): a match that
runs into one keeps playing with the offending spawn / build dropped and the status bit set -- the C oracle
(oracle/lux_oracle.c spawn_unit / spawn_city_tile) defines that behaviour, the kernels must reproduce it byte for
byte, and the other matches of the same warp / CTA / batch must be unaffected.  Several units stacked on one
NON-city cell (a city fell under them and they survived the night, game.py:1173) are inside the engine's domain:
the C oracle is first shown to equal the unmodified reference on such a state (needs /root/reference), then the
kernels to equal the oracle.

Cases run on the host emulation of the kernels here and through the C ABI on the GPU (-m gpu)."""
import numpy as np
import pytest

from luxpythonenvgym_b200 import state as S
from luxpythonenvgym_b200 import wire
from oracle import coracle, stream
from tests import golden_util as G

SIZE = 32
CAPS = S.default_caps(SIZE)          # units 252, cities 255, tiles 384, res 192
SEED = 777


def build(lines, turn=3, uid=1000, cid=1000):
    ms = wire.state_from_updates(SIZE, lines + ["D_DONE"], turn=turn, unit_id_count=uid, city_id_count=cid, caps=CAPS,
                                 recount_adjacency=True)
    assert ms.mirror_errors() == []
    return ms


def units_overflow_state():
    """252 units (the capacity) and two 130-tile cities: either team may still spawn by the game's rule
    (126 units < 130 tiles), the packed state cannot hold a 253rd unit."""
    lines = ["rp 0 0", "rp 1 0", "c 0 c_1 100000 0", "c 1 c_2 100000 0"]
    tiles0 = [(x, y) for y in range(0, 5) for x in range(32)][:130]
    tiles1 = [(x, y) for y in range(20, 25) for x in range(32)][:130]
    lines += ["ct 0 c_1 %d %d 0" % p for p in tiles0] + ["ct 1 c_2 %d %d 0" % p for p in tiles1]
    uid = 1
    for team, tiles in ((0, tiles0), (1, tiles1)):
        for k in range(126):
            x, y = tiles[k]
            lines.append("u %d %d u_%d %d %d 0 0 0 0" % (k % 2 if k < 20 else 0, team, uid, x, y))
            uid += 1
    ms = build(lines, uid=uid - 1, cid=2)
    assert int(ms.hdr["n_units"]) == 252
    tb = np.zeros(CAPS["tiles"], dtype=np.uint8)
    tb[129] = S.TA_BUILD_WORKER               # team 0's last tile: the 253rd unit
    tb[130] = S.TA_RESEARCH
    return ms, np.zeros(CAPS["units"], dtype=np.uint32), tb, 1      # LUX_ST_UNIT_OVERFLOW


def cities_overflow_state():
    """255 one-tile cities (the capacity) and a loaded worker founding the 256th."""
    lines = ["rp 0 0", "rp 1 0"]
    spots = [(2 * i, 2 * j) for j in range(16) for i in range(16) if (i, j) != (15, 15)]
    for k, (x, y) in enumerate(spots):
        lines.append("c %d c_%d 5000 23" % (k % 2, k + 1))
        lines.append("ct %d c_%d %d %d 0" % (k % 2, k + 1, x, y))
    lines.append("u 0 0 u_1 31 31 0 100 0 0")
    lines.append("u 0 1 u_2 29 31 0 60 4 0")
    ms = build(lines, uid=2, cid=255)
    assert int(ms.hdr["n_cities"]) == 255
    uw = np.zeros(CAPS["units"], dtype=np.uint32)
    uw[0] = S.UA_BCITY
    return ms, uw, np.zeros(CAPS["tiles"], dtype=np.uint8), 2      # LUX_ST_CITY_OVERFLOW


def tiles_overflow_state():
    """One 384-tile city (the tile capacity) and a loaded worker building next to it."""
    lines = ["rp 0 0", "rp 1 0", "c 0 c_1 1000000 0"]
    tiles = [(x, y) for y in range(12) for x in range(32)]
    lines += ["ct 0 c_1 %d %d 0" % p for p in tiles]
    lines.append("u 0 0 u_1 5 12 0 100 0 0")
    lines.append("u 0 1 u_2 20 20 0 100 0 0")
    ms = build(lines, uid=2, cid=1)
    assert int(ms.hdr["n_tiles"]) == 384
    uw = np.zeros(CAPS["units"], dtype=np.uint32)
    uw[0] = S.UA_BCITY
    uw[1] = S.UA_BCITY                        # team 1 founds a city far away: same overflow
    return ms, uw, np.zeros(CAPS["tiles"], dtype=np.uint8), 4      # LUX_ST_TILE_OVERFLOW


def stacked_lines():
    """Three team-0 workers and a cart on ONE open cell next to wood and coal, a team-1 worker stacked with a
    team-1 cart on another cell touching the same wood, and bystanders; both teams may mine coal."""
    return ["rp 0 60", "rp 1 55",
            "r wood 10 10 37", "r wood 11 11 400", "r coal 9 11 30", "r wood 12 10 9", "r uranium 20 20 300",
            "u 0 0 u_1 10 11 0 0 0 0", "u 0 0 u_2 10 11 0 90 0 0", "u 0 0 u_3 10 11 2 10 5 0", "u 1 0 u_4 10 11 0 0 0 0",
            "u 0 1 u_5 11 10 0 0 0 0", "u 1 1 u_6 11 10 0 100 0 0", "u 0 1 u_7 12 11 0 99 0 0",
            "u 0 0 u_8 3 3 0 100 0 0",
            "c 0 c_1 400 23", "ct 0 c_1 3 4 0", "c 1 c_2 400 23", "ct 1 c_2 28 28 0", "ccd 10 11 2.5"]


def stacked_carts_lines(road):
    """Two team-0 carts and a team-1 cart on ONE open cell (a stack), road level `road` under them; bystanders."""
    return ["rp 0 0", "rp 1 0", "r wood 20 20 400",
            "u 1 0 u_1 10 11 0 0 0 0", "u 1 0 u_2 10 11 0 0 0 0", "u 1 1 u_3 10 11 0 0 0 0",
            "u 0 0 u_4 5 5 0 0 0 0", "u 0 1 u_5 6 6 0 0 0 0", "u 1 1 u_6 15 15 0 0 0 0",
            "c 0 c_1 400 23", "ct 0 c_1 3 4 0", "c 1 c_2 400 23", "ct 1 c_2 28 28 0"] + \
           (["ccd 10 11 %s" % road] if road else [])


def normal_states(k):
    z = G.npz("dense")
    names = [n for n in G.dense_names() if int(z[n + "/size"]) == SIZE]
    out = []
    for i in range(k):
        ms = G.get_state(z, names[i % len(names)] + "/init")
        o = S.MatchState(SIZE, CAPS)
        o.hdr, o.cells = ms.hdr.copy(), ms.cells.copy()
        for sec in ("units", "cities", "tiles", "res"):
            dst, src = getattr(o, sec), getattr(ms, sec)
            n = min(len(dst), len(src))
            dst[:n] = src[:n]
        out.append(o)
    return out


def special_cases():
    cases = [units_overflow_state(), cities_overflow_state(), tiles_overflow_state()]
    stack = build(stacked_lines(), uid=8, cid=2)
    cases.append((stack, np.zeros(CAPS["units"], dtype=np.uint32), np.zeros(CAPS["tiles"], dtype=np.uint8), 0))
    return cases


def batch_of_states():
    """8 matches: the four special ones interleaved with four ordinary dense ones (so that every special match
    shares its warp pair / CTA with an ordinary neighbour)."""
    special = special_cases()
    normal = normal_states(4)
    states, first_u, first_t, expect = [], [], [], []
    for i in range(4):
        states += [special[i][0], normal[i]]
        first_u += [special[i][1], None]
        first_t += [special[i][2], None]
        expect += [special[i][3], 0]
    return states, first_u, first_t, expect


def oracle_play(states, first_u, first_t, turns):
    """The oracle's trajectory: per turn a list of MatchState copies, and the action words used."""
    cur = [s.copy() for s in states]
    traj, acts = [], []
    for t in range(turns):
        uws, tbs = [], []
        for m, ms in enumerate(cur):
            uw, tb = coracle.gen_actions(ms, SEED, m)
            if t == 0 and first_u[m] is not None:
                uw, tb = first_u[m].copy(), first_t[m].copy()
            uws.append(uw)
            tbs.append(tb)
            coracle.step(ms, uw, tb)
        acts.append((uws, tbs))
        traj.append([s.copy() for s in cur])
    return traj, acts


@pytest.mark.parametrize("group", [32, 8])
def test_overflow_and_stack_states_on_the_emulated_kernels(group):
    from tests.emu import emu
    states, first_u, first_t, expect = batch_of_states()
    traj, acts = oracle_play(states, first_u, first_t, 6)
    cur = [s.copy() for s in states]
    for t in range(6):
        emu.step_many(cur, acts[t][0], acts[t][1], group=group)
        for m in range(len(cur)):
            assert traj[t][m].diff(cur[m]) == [], (t, m)
            assert int(cur[m].hdr["status"]) == int(traj[t][m].hdr["status"])
    for m, want in enumerate(expect):
        assert int(cur[m].hdr["status"]) == want, (m, int(cur[m].hdr["status"]))
    assert int(cur[0].hdr["n_units"]) <= 252 and int(cur[2].hdr["n_cities"]) <= 255 and int(cur[4].hdr["n_tiles"]) <= 384


@pytest.mark.gpu
@pytest.mark.parametrize("two_class,group", [(-1, 0), (0, 0), (1, 16)])
def test_overflow_and_stack_states_on_the_gpu(two_class, group):
    import torch
    from luxpythonenvgym_b200 import _lib
    from luxpythonenvgym_b200.batched import BatchedGame
    lib = _lib.load()
    states, first_u, first_t, expect = batch_of_states()
    # 64 matches: the 8-match pattern repeated, so that the special matches meet every warp / CTA position
    reps = 8
    all_states = states * reps
    traj, acts = oracle_play(all_states, first_u * reps, first_t * reps, 6)
    try:
        lib.lux_b200_set_option(b"two_class", two_class)
        lib.lux_b200_set_option(b"group", group)
        game = BatchedGame(len(all_states), SIZE, caps=CAPS)
        game.load_states(all_states)
        for t in range(6):
            ua = np.stack(acts[t][0]).view(np.int32)
            ta = np.stack(acts[t][1])
            game.unit_actions.copy_(torch.from_numpy(ua))
            game.tile_actions.copy_(torch.from_numpy(ta))
            game.step()
            got = game.export_states()
            for m in range(len(all_states)):
                assert traj[t][m].diff(got[m]) == [], (t, m)
                assert int(got[m].hdr["status"]) == int(traj[t][m].hdr["status"]), (t, m)
        for m in range(len(all_states)):
            assert int(got[m].hdr["status"]) == expect[m % 8], m
    finally:
        lib.lux_b200_set_option(b"two_class", -1)
        lib.lux_b200_set_option(b"group", 0)


@pytest.mark.gpu
@pytest.mark.parametrize("two_class,tma", [(-1, 1), (0, 1), (0, 0)])
def test_unplayable_headers_are_flagged_and_left_alone(two_class, tma):
    """A caller-built state whose header cannot be played safely -- per-team unit counts that do not add up, a live
    count beyond the batch capacity -- is left untouched with LUX_ST_BAD_STATE set; its neighbours play on, equal
    to the oracle.  Two-class step (the classification pass checks), single launch, plain staging."""
    import torch
    from luxpythonenvgym_b200 import _lib
    from luxpythonenvgym_b200.batched import BatchedGame
    lib = _lib.load()
    states = normal_states(16)
    bad_a, bad_b = 3, 10
    states[bad_a].hdr["n_team_units"][0] += 1                      # 2 + 1 != n_units
    states[bad_b].hdr["n_tiles"] = CAPS["tiles"] + 16              # beyond t_cap
    good = [m for m in range(16) if m not in (bad_a, bad_b)]
    traj, acts = oracle_play([states[m] for m in good], [None] * len(good), [None] * len(good), 5)
    try:
        lib.lux_b200_set_option(b"two_class", two_class)
        lib.lux_b200_set_option(b"tma", tma)
        game = BatchedGame(16, SIZE, caps=CAPS)
        game.load_states(states)
        before = game.export_states()
        for t in range(5):
            ua = np.zeros((16, CAPS["units"]), dtype=np.uint32)
            ta = np.zeros((16, CAPS["tiles"]), dtype=np.uint8)
            for j, m in enumerate(good):
                ua[m], ta[m] = acts[t][0][j], acts[t][1][j]
            game.unit_actions.copy_(torch.from_numpy(ua.view(np.int32)))
            game.tile_actions.copy_(torch.from_numpy(ta))
            game.step()
            got = game.export_states()
            for j, m in enumerate(good):
                assert traj[t][j].diff(got[m]) == [], (t, m)
            for m in (bad_a, bad_b):
                assert int(got[m].hdr["status"]) & S.ST_BAD_STATE, (t, m)
                assert int(got[m].hdr["turn"]) == int(before[m].hdr["turn"]) and int(got[m].hdr["done"]) == 0
                assert np.array_equal(got[m].cells, before[m].cells) and np.array_equal(got[m].units, before[m].units)
    finally:
        lib.lux_b200_set_option(b"two_class", -1)
        lib.lux_b200_set_option(b"tma", 1)


@pytest.mark.gpu
@pytest.mark.parametrize("road", [0, 4.5, 5.25])
def test_stacked_carts_build_their_road_in_turn_on_the_gpu(road):
    """Several carts of a stack each add their road step in turn (unit.py:257-269), the later ones only while the road
    is below the maximum -- which also decides whose statistics the step goes to.  The lanes of the unit pass all
    start from the level they load, so the stack path redoes the steps in unit order; idle turns (the carts just
    stand there), from an empty road and from levels where the maximum cuts the sequence short."""
    import torch
    from luxpythonenvgym_b200.batched import BatchedGame
    ms = build(stacked_carts_lines(road), uid=8, cid=2)
    others = normal_states(3)
    states = [others[0], ms, others[1], ms.copy(), others[2]]
    traj = [s.copy() for s in states]
    game = BatchedGame(len(states), SIZE, caps=CAPS)
    game.load_states(states)
    zu, zt = np.zeros(CAPS["units"], dtype=np.uint32), np.zeros(CAPS["tiles"], dtype=np.uint8)
    cell = 11 * SIZE + 10
    for t in range(6):
        for m in traj:
            coracle.step(m, zu, zt)
        game.unit_actions.zero_()
        game.tile_actions.zero_()
        game.step()
        got = game.export_states()
        for m in range(len(states)):
            assert traj[m].diff(got[m]) == [], (t, m)
    assert int(traj[1].cells[cell]["road_q"]) == 24 and int(traj[1].hdr["status"]) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 5])
def test_random_stacks_soak_on_the_gpu(seed):
    """scripts/stack_soak.py as a test: random hand-built states with stacks of both teams' workers and carts on open
    cells next to resources and on roads, 25 turns of the canonical action stream, GPU against the C oracle on every
    turn.  This is where the orders of a stack's road work (pillage / cart steps / a city founded under earlier
    units) were found to matter; a match in which a second unit founds a city on a cell an earlier one has just built
    on is flagged LUX_ST_BAD_STATE and left alone (a handful per few hundred states)."""
    import subprocess, sys, os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "stack_soak.py"), "160", "25", str(seed)],
                         capture_output=True, text=True, cwd=root, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"mismatching match-turns: (\d+), matches flagged LUX_ST_BAD_STATE \(left alone\): (\d+)", out.stdout)
    assert m, out.stdout[-2000:]
    assert int(m.group(1)) == 0, out.stdout[-3000:]
    assert int(m.group(2)) <= 8


@pytest.mark.gpu
def test_nightfall_chaos_soak_on_the_gpu():
    """scripts/chaos_soak.py as a test: random hand-built states at nightfall -- multi-tile cities of both teams with
    little fuel and several units on their tiles, cargo that does or does not last the night, resources and roads
    around -- 45 turns of the canonical action stream, GPU against the C oracle on every turn (city falls under
    units, starvation, merges of cities built next to each other)."""
    import subprocess, sys, os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "chaos_soak.py"), "160", "45", "7"],
                         capture_output=True, text=True, cwd=root, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"mismatching match-turns: (\d+), flagged LUX_ST_BAD_STATE: (\d+)", out.stdout)
    assert m, out.stdout[-2000:]
    assert int(m.group(1)) == 0 and int(m.group(2)) == 0, out.stdout[-3000:]


@pytest.mark.reference
def test_oracle_equals_reference_on_stacked_carts():
    """The unmodified reference on the stacked-carts state: cart after cart adds its road step; the C oracle
    reproduces every turn (idle turns)."""
    import os
    from oracle import flatten, refshim
    ref = refshim.load()
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        for road in (0, 4.5, 5.25):
            cfg = dict(ref.LuxMatchConfigs_Default)
            cfg.update(seed=None, mapType="empty", width=SIZE, height=SIZE)
            game = ref.Game(cfg)
            game.log = lambda text: None
            game.reset(updates=stacked_carts_lines(road) + ["D_DONE"])
            game.global_unit_id_count, game.global_city_id_count = 8, 2
            game.state["turn"] = 3
            mine = flatten.flatten(game, caps=CAPS)
            assert len(game.map.get_cell(10, 11).units) == 3
            zu, zt = np.zeros(CAPS["units"], dtype=np.uint32), np.zeros(CAPS["tiles"], dtype=np.uint8)
            for turn in range(6):
                game.run_turn_with_actions([])
                coracle.step(mine, zu, zt)
                assert flatten.flatten(game, caps=CAPS).diff(mine) == [], (road, turn)
            assert game.map.get_cell(10, 11).get_road() == 6.0
    finally:
        os.chdir(cwd)


@pytest.mark.reference
def test_oracle_equals_reference_on_stacked_units():
    """The unmodified reference plays the stacked state (imported with its own process_updates) for 25 turns of
    the canonical action stream; the C oracle, started from the flattened state, reproduces every turn."""
    import os
    from oracle import flatten, refshim
    ref = refshim.load()
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        cfg = dict(ref.LuxMatchConfigs_Default)
        cfg.update(seed=None, mapType="empty", width=SIZE, height=SIZE)
        game = ref.Game(cfg)
        game.log = lambda text: None
        game.reset(updates=stacked_lines() + ["D_DONE"])
        game.global_unit_id_count, game.global_city_id_count = 8, 2
        game.state["turn"] = 3
        ms = flatten.flatten(game, caps=CAPS)
        mine = ms.copy()
        assert len(game.map.get_cell(10, 11).units) == 4 and not game.map.get_cell(10, 11).is_city_tile()
        mined = 0
        for turn in range(25):
            uw, tb = stream.gen_actions(ms, SEED, 0)
            before = sum(u.cargo["wood"] + u.cargo["coal"] for u in game.get_teams_units(0).values())
            over = game.run_turn_with_actions(flatten.words_to_actions(ref, game, uw, tb))
            mined += sum(u.cargo["wood"] + u.cargo["coal"] for u in game.get_teams_units(0).values()) != before
            coracle.step(mine, uw, tb)
            ms = flatten.flatten(game, caps=CAPS)
            if over:
                ms.hdr["done"], ms.hdr["winner"] = 1, flatten.winner_code(game)
            assert ms.diff(mine) == [], turn
            if over:
                break
        assert mined > 0
    finally:
        os.chdir(cwd)
