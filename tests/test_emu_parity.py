"""CUDA kernel LOGIC on the CPU: luxpythonenvgym_b200/csrc/*.cuh compiled for the host with a
32-fiber warp emulator (tests/emu/) and compared with the oracle / golden vectors.  This is a
development aid for a container without a GPU -- the product never runs this build; the real
parity tests are tests/test_gpu_parity.py (-m gpu)."""
import numpy as np
import pytest

from luxpythonenvgym_b200 import state as S
from oracle import coracle
from tests import golden_util as G
from tests.emu import emu

CAPS = dict(units=252, cities=255, tiles=512, res=384)


def recap(ms):
    o = S.MatchState(ms.size, CAPS)
    o.hdr, o.cells = ms.hdr.copy(), ms.cells.copy()
    for name in ("units", "cities", "tiles", "res"):
        dst, src = getattr(o, name), getattr(ms, name)
        k = min(len(dst), len(src))
        dst[:k] = src[:k]
    return o


@pytest.mark.parametrize("rid", G.replay_ids()[:3])
def test_emulated_step_replays(rid):
    z = G.npz("replays")
    init = G.get_state(z, rid + "/init")
    U, T = G.replay_actions(z, rid, init)
    ms = recap(init)
    digests = z[rid + "/digests"]
    for turn in range(len(digests)):
        emu.step(ms, U[turn], T[turn])
        assert ms.digest() == int(digests[turn]), "replay %s turn %d" % (rid, turn)


@pytest.mark.parametrize("name", G.dense_names())
def test_emulated_step_dense(name):
    z = G.npz("dense")
    seed, start = int(z["stream_seed"]), int(name.split("@")[1])
    a = recap(G.get_state(z, name + "/init"))
    b = a.copy()
    digests = z[name + "/digests"]
    for turn in range(len(digests)):
        uw, tb = coracle.gen_actions(a, seed, start)
        coracle.step(a, uw, tb)
        emu.step(b, uw, tb)
        assert a.diff(b) == []
        assert b.digest() == int(digests[turn])


@pytest.mark.parametrize("k", range(0, 32, 3))
def test_emulated_step_random(k):
    z = G.npz("random")
    seed = int(z["stream_seed"])
    ms = recap(G.get_state(z, "%d/init" % k))
    digests = z["%d/digests" % k]
    for turn in range(len(digests)):
        uw, tb = coracle.gen_actions(ms, seed, k)
        emu.step(ms, uw, tb)
        assert ms.digest() == int(digests[turn]), "match %d turn %d" % (k, turn)


@pytest.mark.parametrize("group", [8, 16])
def test_emulated_step_lockstep_groups(group):
    """Several different matches share one warp (32 / group lane groups in lockstep): dense late-game
    states next to fresh sparse ones, different turn counters (day and night in one warp), a finished
    match and an odd batch size so that idle groups walk along."""
    zd, zr = G.npz("dense"), G.npz("random")
    names = [nm for nm in G.dense_names() if int(zd[nm + "/size"]) == 32][:3]
    a = [recap(G.get_state(zd, nm + "/init")) for nm in names]
    ids = [int(nm.split("@")[1]) for nm in names]
    for k in range(int(zr["n"])):
        ms = recap(G.get_state(zr, "%d/init" % k))
        if ms.size == 32 and len(a) < 9:
            a.append(ms)
            ids.append(k)
    assert len(a) >= 6
    b = [ms.copy() for ms in a]
    seed = int(zd["stream_seed"])
    for turn in range(70):
        acts = [coracle.gen_actions(ms, seed, i) for ms, i in zip(a, ids)]
        for ms, (uw, tb) in zip(a, acts):
            if not ms.hdr["done"]:
                coracle.step(ms, uw, tb)
        emu.step_many(b, [x[0] for x in acts], [x[1] for x in acts], group=group)
        for i, (x, y) in enumerate(zip(a, b)):
            assert x.diff(y) == [], "group %d turn %d match %d" % (group, turn, i)


@pytest.mark.parametrize("name", G.dense_names()[::3])
def test_turn_need_bounds_the_turn_and_slices_are_respected(name):
    """The per-match shared-memory layout is sized by what the header says the turn can touch (LuxNeed).
    Property, turn by turn on dense late-game states: (1) the populations after the turn never exceed the
    need computed before it; (2) played with a slice of exactly that many bytes the emulated kernel matches
    the oracle and never writes behind the slice (canary); (3) one byte less and the match is left untouched
    for the big class."""
    z = G.npz("dense")
    seed, start = int(z["stream_seed"]), int(name.split("@")[1])
    a = recap(G.get_state(z, name + "/init"))
    b = a.copy()
    for turn in range(30):
        u, c, t, r, q, total = emu.need(a)
        uw, tb = coracle.gen_actions(a, seed, start)
        if turn % 7 == 3:
            before = b.copy()
            assert emu.step_slice(b, uw, tb, total - 16) is True
            assert before.diff(b) == []
        coracle.step(a, uw, tb)
        assert emu.step_slice(b, uw, tb, total) is False
        assert a.diff(b) == [], "turn %d" % turn
        assert int(a.hdr["n_units"]) <= u and int(a.hdr["n_cities"]) <= c and int(a.hdr["n_tiles"]) <= t
        assert int(a.hdr["status"]) == 0


@pytest.mark.parametrize("name", G.obs_names())
def test_emulated_observations(name):
    z = G.npz("obs")
    ms = recap(G.get_state(z, name + "/state"))
    for team in (0, 1):
        want, codes = z["%s/obs%d" % (name, team)], z["%s/ent%d" % (name, team)]
        got, got_codes = emu.observe(ms, team)
        assert np.array_equal(codes, got_codes)
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), np.argwhere(want != got)[:5]


@pytest.mark.parametrize("k", [0, 7, 10, 12, 16, 19])
def test_emulated_env_flow(k):
    """Pre-validation (LUX_STEP_PREVALIDATE) inside the emulated step kernel + emulated observations
    on the LuxEnvironment fixtures; action decoding / reward come from the oracle here."""
    from tests import env_util

    def step_turn(ms, ent, codes):
        uw, tb = coracle.encode_actions(ms, ent, codes)
        emu.step(ms, uw, tb, flags=1)

    ms = recap(G.get_state(G.npz("env"), "%d/init" % k))
    env_util.check_episode(k, ms, emu.observe, step_turn, coracle.reward)
