"""Helpers for the -m gpu parity tests: batched oracle on host copies, vectorised diff."""
import ctypes

import numpy as np

from luxpythonenvgym_b200 import state as S
from oracle import coracle


class HostBatch:
    """Host mirror of a BatchedGame driven by the C oracle (the checker)."""

    def __init__(self, game):
        self.size, self.caps, self.n = game.size, game.caps, game.n
        self.arr = {k: np.ascontiguousarray(t.cpu().numpy()) for k, t in game.sections().items()}
        self.desc = coracle.batch_desc(self.size, self.caps, self.arr, self.n)
        self.ua = np.zeros((self.n, self.caps["units"]), dtype=np.uint32)
        self.ta = np.zeros((self.n, self.caps["tiles"]), dtype=np.uint8)

    def gen_actions(self, seed, base=0):
        coracle.gen_actions_batch(self.desc, seed, base, self.ua, self.ta)

    def step(self, ua=None, ta=None):
        coracle.step_batch(self.desc, self.ua if ua is None else ua, self.ta if ta is None else ta)

    def hdr(self):
        return self.arr["hdr"].view(S.header_dtype).reshape(self.n)


def diff_batches(host, game):
    """Indices of matches whose live state differs between the oracle copy and the GPU."""
    dev = {k: t.cpu().numpy() for k, t in game.sections().items()}
    hh = host.hdr()
    bad = np.zeros(host.n, dtype=bool)
    bad |= (host.arr["hdr"] != dev["hdr"]).any(axis=1)
    bad |= (host.arr["cells"] != dev["cells"]).any(axis=1)
    for name, rec, count in (("units", 16, hh["n_units"]), ("cities", 16, hh["n_cities"]), ("tiles", 2, hh["n_tiles"]),
                             ("res", 4, hh["n_res"])):
        cap = host.arr[name].shape[1] // rec
        live = (np.arange(cap)[None, :] < count[:, None]).repeat(rec, axis=1)
        bad |= ((host.arr[name] != dev[name]) & live).any(axis=1)
    return np.nonzero(bad)[0]


def explain(host, game, m):
    a = S.MatchState.from_arrays(host.size, {k: v[m] for k, v in host.arr.items()})
    b = game.export_states([m])[0]
    return a.diff(b)
