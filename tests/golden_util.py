"""Loaders for tests/golden/*.npz (produced by oracle/make_golden.py from the reference)."""
import os

import numpy as np

from luxpythonenvgym_b200 import state as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def npz(name):
    if name not in _cache:
        _cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
    return _cache[name]


def get_state(z, prefix):
    size = int(z[prefix.split("/")[0] + "/size"])
    # the fixtures hold the resource list as uint16 cell indices (round-1 layout); digests are defined over that
    # view (MatchState.live), so they stay valid for the LuxRes layout
    return S.MatchState.from_arrays(size, {k: z["%s/%s" % (prefix, k)] for k in S.MatchState.SECTIONS}, legacy_res=True)


def replay_ids():
    return [str(x) for x in npz("replays")["ids"]]


def replay_actions(z, rid, ms):
    """Per-turn dense action tensors of a replay fixture."""
    ua, ta = z[rid + "/unit_actions"], z[rid + "/tile_actions"]
    turns = len(z[rid + "/digests"])
    U = np.zeros((turns, len(ms.units)), dtype=np.uint32)
    T = np.zeros((turns, len(ms.tiles)), dtype=np.uint8)
    U[ua[:, 0], ua[:, 1]] = ua[:, 2]
    T[ta[:, 0], ta[:, 1]] = ta[:, 2].astype(np.uint8)
    return U, T


def dense_names():
    return [str(x) for x in npz("dense")["names"]]


def obs_names():
    return [str(x) for x in npz("obs")["names"]]
