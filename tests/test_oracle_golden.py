"""The C oracle (oracle/lux_oracle.c) against golden vectors made by the unmodified reference.

Bit-exact bar: every per-turn digest of the canonical packed state must match, and the
stored full states must be identical field by field.
"""
import numpy as np
import pytest

from oracle import coracle, stream
from tests import golden_util as G


@pytest.mark.parametrize("rid", G.replay_ids())
def test_oracle_replays(rid):
    z = G.npz("replays")
    ms = G.get_state(z, rid + "/init")
    U, T = G.replay_actions(z, rid, ms)
    digests = z[rid + "/digests"]
    for turn in range(len(digests)):
        over = coracle.step(ms, U[turn], T[turn])
        assert ms.digest() == int(digests[turn]), "replay %s diverged at turn %d" % (rid, turn)
        key = "%s/t%d" % (rid, turn + 1)
        if key + "/hdr" in z.files:
            assert G.get_state(z, key).diff(ms) == []
        assert over == (turn == len(digests) - 1)


@pytest.mark.parametrize("k", range(int(G.npz("random")["n"])))
def test_oracle_random_stream(k):
    z = G.npz("random")
    seed = int(z["stream_seed"])
    ms = G.get_state(z, "%d/init" % k)
    digests = z["%d/digests" % k]
    for turn in range(len(digests)):
        uw, tb = coracle.gen_actions(ms, seed, k)
        if turn % 16 == 0:  # the Python and C statements of the stream agree
            uw2, tb2 = stream.gen_actions(ms, seed, k)
            assert np.array_equal(uw, uw2) and np.array_equal(tb, tb2)
        coracle.step(ms, uw, tb)
        assert ms.digest() == int(digests[turn]), "match %d diverged at turn %d" % (k, turn)
    assert G.get_state(z, "%d/final" % k).diff(ms) == []


@pytest.mark.parametrize("name", G.dense_names())
def test_oracle_dense(name):
    z = G.npz("dense")
    seed = int(z["stream_seed"])
    start = int(name.split("@")[1])
    ms = G.get_state(z, name + "/init")
    digests = z[name + "/digests"]
    for turn in range(len(digests)):
        uw, tb = coracle.gen_actions(ms, seed, start)
        coracle.step(ms, uw, tb)
        assert ms.digest() == int(digests[turn]), "%s diverged at step %d" % (name, turn)
    assert G.get_state(z, name + "/final").diff(ms) == []


def test_oracle_done_is_sticky():
    z = G.npz("random")
    ms = G.get_state(z, "0/final")
    assert ms.hdr["done"] == 1
    before = ms.digest()
    coracle.step(ms, None, None)
    assert ms.digest() == before


def test_wood_regrowth_table():
    """ceil(min(a*1.025, 500)) == min((41a+39)//40, 500) on 0..499."""
    import math
    for a in range(0, 500):
        assert math.ceil(min(a * 1.025, 500)) == min((41 * a + 39) // 40, 500)
