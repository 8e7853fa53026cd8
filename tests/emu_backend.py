"""TEST INFRASTRUCTURE ONLY -- a stand-in for `BatchedGame` that runs the CUDA kernels' source on the host.

The single-match adapters (`Game`, `MatchController`, `LuxEnvironment`, the agents) are host protocol code over
`BatchedGame`.  To test that protocol code in the CPU suite, `use_emulated_batch(monkeypatch)` swaps the class
the adapters instantiate for `EmuBatchedGame`, which keeps the match as a host `MatchState` and plays turns /
builds observation rows through tests/emu (csrc/lux_step_core.cuh and lux_obs_core.cuh compiled for the host
with the warp emulator).  The `-m gpu` tests run the same cases on the real `BatchedGame`.  Nothing under
luxpythonenvgym_b200/ imports this module; without the patch the adapters raise when there is no CUDA device.
"""
import numpy as np
import torch

from luxpythonenvgym_b200 import state as S
from tests.emu import emu


class EmuBatchedGame:
    def __init__(self, n_matches, size, caps=None, device="cpu"):
        assert int(n_matches) == 1, "the emulated stand-in holds one match"
        self.n, self.size = 1, int(size)
        self.caps = dict(S.default_caps(size), **(caps or {}))
        self.device = torch.device("cpu")
        self._ms = S.MatchState(self.size, self.caps)
        self._ms.hdr["done"] = 1
        self._uw = np.zeros(self.caps["units"], dtype=np.uint32)
        self._tb = np.zeros(self.caps["tiles"], dtype=np.uint8)
        self.steps = 0

    def load_states(self, states, first=0):
        assert len(states) == 1 and first == 0
        ms = S.MatchState(self.size, self.caps)
        src = states[0]
        ms.hdr, ms.cells = src.hdr.copy(), src.cells.copy()
        for name in ("units", "cities", "tiles", "res"):
            a, b = getattr(src, name), getattr(ms, name)
            b[: len(a)] = a[: len(b)]
        self._ms = ms

    def export_states(self, indices=None):
        return [self._ms.copy()]

    def put_actions(self, m, unit_words, tile_bytes):
        self._uw[:] = 0
        self._tb[:] = 0
        self._uw[: len(unit_words)] = unit_words
        self._tb[: len(tile_bytes)] = tile_bytes

    def step(self, unit_actions=None, tile_actions=None, prevalidate=False):
        if not self._ms.hdr["done"]:
            emu.step(self._ms, self._uw, self._tb, flags=1 if prevalidate else 0)
        self.steps += 1

    def observe(self, team_sel, e_cap=None):
        e_cap = int(e_cap or (self.caps["units"] + self.caps["tiles"]))
        rows, ent = emu.observe(self._ms, int(team_sel), e_cap)
        obs = torch.zeros((1, e_cap, 85), dtype=torch.float32)
        codes = torch.zeros((1, e_cap), dtype=torch.int16)
        obs[0, : len(rows)] = torch.from_numpy(rows)
        codes[0, : len(ent)] = torch.from_numpy(ent)
        return obs, codes, torch.tensor([len(rows)], dtype=torch.int32)

    def unit_word(self, m, slot, code):
        return emu.unit_word(self._ms, slot, code)


def use_emulated_batch(monkeypatch):
    import luxpythonenvgym_b200.game as game_module
    monkeypatch.setattr(game_module, "BatchedGame", EmuBatchedGame)
