"""Shared driver for the LuxEnvironment-flow fixtures (tests/golden/env.npz)."""
import numpy as np

from oracle import stream
from tests import golden_util as G


def env_ids():
    return list(range(int(G.npz("env")["n"])))


def policy_codes(seed, k, turn, ms, ent):
    """The scripted policy of oracle/make_golden.make_env: hash of (match, turn, entity) % 9."""
    codes = np.zeros(len(ent), dtype=np.int32)
    for i, e in enumerate(ent):
        e = int(e)
        if e & 0x4000:
            kind, key = 3, int(ms.tiles[e & 0x3FFF])
        else:
            kind, key = 2, int(ms.units["id"][e])
        codes[i] = stream.hash64(seed, k, turn, kind, key) % 9
    return codes


def check_episode(k, ms, observe, step_turn, reward):
    """Runs fixture k.  observe(ms, team) -> (obs, ent); step_turn(ms, ent, codes) advances one turn with
    pre-validation; reward(ms, team, last) -> float.  Asserts decisions, digests, observations, rewards."""
    z = G.npz("env")
    seed, team = int(z["stream_seed"]), int(z["%d/team" % k])
    dec, digests, want_obs = z["%d/decisions" % k], z["%d/digests" % k], z["%d/obs" % k]
    last = np.zeros(3, dtype=np.int32)
    di, oi, total, ref_total = 0, 0, 0.0, 0.0
    for turn_idx in range(len(digests)):
        turn = int(ms.hdr["turn"])
        obs, ent = observe(ms, team)
        codes = policy_codes(seed, k, turn, ms, ent)
        rows = dec[di:di + len(ent)]
        assert len(rows) == len(ent) and (rows[:, 0] == turn).all(), "turn %d: entity count" % turn
        assert np.array_equal(rows[:, 1].astype(np.int64), np.asarray(ent, dtype=np.int64)), "turn %d: entity order" % turn
        assert np.array_equal(rows[:, 2].astype(np.int64), codes), "turn %d" % turn
        if turn < 50 and len(want_obs):
            n = len(ent)
            assert np.array_equal(want_obs[oi:oi + n].view(np.uint32), np.asarray(obs, dtype=np.float32).view(np.uint32))
            oi += n
        # the rewards the reference returned for all earlier decisions cover exactly the turns played so far
        if len(rows):
            assert abs(ref_total - total) < 1e-9, "turn %d: reward %r vs %r" % (turn, ref_total, total)
        ref_total += float(rows[:, 3].sum())
        di += len(ent)
        step_turn(ms, ent, codes)
        assert ms.digest() == int(digests[turn_idx]), "fixture %d diverged at turn %d" % (k, turn)
        total += reward(ms, team, last)
    if di < len(dec):      # the reward of the final decision (game end) closes the episode
        raise AssertionError("unconsumed decisions")
    if len(dec) and dec[-1, 4] == 1.0:
        assert abs(float(dec[:, 3].sum()) - total) < 1e-9
