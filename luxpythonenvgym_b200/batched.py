"""BatchedGame: thousands of independent Lux AI 2021 matches resident on one GPU.

Host-side mirror of the reference's `Game` for the per-turn path
(luxai2021/game/game.py:390 `run_turn_with_actions`), batched: `step(unit_actions,
tile_actions)` advances every live match by one turn with one kernel launch.  State lives
in six torch uint8 tensors (one per section of luxpythonenvgym_b200/state.py); torch is
only the allocator / stream provider, all compute is in _lux_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import state as S

# torch's current stream as a raw handle (an int ctypes passes as void*); the public accessor builds a Stream object
_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)


class BatchedGame:
    def __init__(self, n_matches, size, caps=None, device="cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedGame needs a CUDA device (B200); there is no CPU fallback")
        self.lib = _lib.load()
        self.n = int(n_matches)
        self.size = int(size)
        self.caps = dict(S.default_caps(size), **(caps or {}))
        self.device = torch.device(device)
        self._dev_index = self.device.index if self.device.index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
        c = self.caps
        ncell = size * size
        z = lambda *shape: torch.zeros(shape, dtype=torch.uint8, device=self.device)
        self.hdr = z(self.n, S.HDR_BYTES)
        self.cells = z(self.n, ncell * 8)
        self.units = z(self.n, c["units"] * 16)
        self.cities = z(self.n, c["cities"] * 16)
        self.tiles = z(self.n, c["tiles"] * 2)
        self.res = z(self.n, c["res"] * 4)
        self.unit_actions = torch.zeros((self.n, c["units"]), dtype=torch.int32, device=self.device)
        self.tile_actions = z(self.n, c["tiles"])
        self._stats = torch.zeros(16, dtype=torch.int64, device=self.device)
        self._desc = self._make_desc()
        # a fresh batch is "done" until states are loaded
        self.hdr[:, 24] = 1

    # ------------------------------------------------------------------ plumbing
    def _make_desc(self):
        c = self.caps
        return _lib.LuxBatch(self.n, self.size, c["units"], c["cities"], c["tiles"], c["res"],
                             self.hdr.data_ptr(), self.cells.data_ptr(), self.units.data_ptr(),
                             self.cities.data_ptr(), self.tiles.data_ptr(), self.res.data_ptr())

    def _stream(self):
        # the raw handle of torch's current stream (no Stream object per call: this sits on the per-turn host path)
        if _RAW_STREAM is not None:
            return _RAW_STREAM(self._dev_index)
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def sections(self):
        return {"hdr": self.hdr, "cells": self.cells, "units": self.units, "cities": self.cities,
                "tiles": self.tiles, "res": self.res}

    # ------------------------------------------------------------------ state import / export
    def load_states(self, states, first=0):
        """Copy a list of MatchState (host) into matches first..first+len-1."""
        k = len(states)
        host = {name: np.zeros((k,) + tuple(t.shape[1:]), dtype=np.uint8) for name, t in self.sections().items()}
        for i, ms in enumerate(states):
            assert ms.size == self.size
            for name, arr in ms.to_bytes().items():
                assert arr.size <= host[name].shape[1], "MatchState %s exceeds batch capacity" % name
                host[name][i, : arr.size] = arr
        for name, t in self.sections().items():
            t[first:first + k].copy_(torch.from_numpy(host[name]))

    def load_arrays(self, arrays, first=0):
        """Copy host section arrays ([k, bytes] uint8, e.g. from mapgen.generate_batch) into matches first.."""
        for name, t in self.sections().items():
            a = arrays[name]
            assert a.shape[1] == t.shape[1], "section %s: capacity mismatch" % name
            t[first:first + a.shape[0]].copy_(torch.from_numpy(a))

    def export_states(self, indices=None):
        """MatchState copies (host) of the given matches (default: all)."""
        idx = list(range(self.n)) if indices is None else list(indices)
        it = torch.as_tensor(idx, device=self.device, dtype=torch.long)
        host = {name: t.index_select(0, it).cpu().numpy() for name, t in self.sections().items()}
        return [S.MatchState.from_arrays(self.size, {name: host[name][i] for name in host}) for i in range(len(idx))]

    def headers(self):
        """Structured numpy view of all headers (one D2H copy)."""
        return self.hdr.cpu().numpy().view(S.header_dtype).reshape(self.n)

    # ------------------------------------------------------------------ the hot path
    def step(self, unit_actions=None, tile_actions=None, prevalidate=False):
        """One turn for every live match.  Actions default to this object's action tensors."""
        ua = self.unit_actions if unit_actions is None else unit_actions
        ta = self.tile_actions if tile_actions is None else tile_actions
        assert ua.is_cuda and ta.is_cuda and ua.is_contiguous() and ta.is_contiguous()
        assert ua.numel() == self.n * self.caps["units"] and ua.element_size() == 4
        assert ta.numel() == self.n * self.caps["tiles"] and ta.element_size() == 1
        _lib.check(self.lib.lux_b200_step(ctypes.byref(self._desc), ua.data_ptr(), ta.data_ptr(),
                                          _lib.LUX_STEP_PREVALIDATE if prevalidate else 0, self._stream()),
                   "lux_b200_step")

    def put_actions(self, m, unit_words, tile_bytes):
        """Host action words of ONE match (numpy uint32 [<= u_cap], uint8 [<= t_cap]) -> the action tensors."""
        uw = np.zeros(self.caps["units"], dtype=np.uint32)
        tb = np.zeros(self.caps["tiles"], dtype=np.uint8)
        uw[: len(unit_words)] = unit_words
        tb[: len(tile_bytes)] = tile_bytes
        self.unit_actions[m].copy_(torch.from_numpy(uw.view(np.int32)))
        self.tile_actions[m].copy_(torch.from_numpy(tb))

    def step_host(self, unit_actions_host, tile_actions_host, hdr_host=None, done_host=None, prevalidate=False,
                  u_used=0, t_used=0):
        """The same turn from HOST (pinned) action buffers; copies back headers/done flags."""
        _lib.check(self.lib.lux_b200_step_host(
            ctypes.byref(self._desc), unit_actions_host.data_ptr(), tile_actions_host.data_ptr(),
            self.unit_actions.data_ptr(), self.tile_actions.data_ptr(),
            done_host.data_ptr() if done_host is not None else None,
            hdr_host.data_ptr() if hdr_host is not None else None, int(u_used), int(t_used),
            _lib.LUX_STEP_PREVALIDATE if prevalidate else 0, self._stream()), "lux_b200_step_host")

    def step_host_sparse(self, entries_host, k, done_host=None, summary_host=None, prevalidate=False):
        """One turn from a host action LIST: entries_host is an int32/uint32 [cap, 2] (pinned) tensor whose
        first k rows are { tile<<31 | match<<10 | slot, action word }; see `pack_entries`."""
        _lib.check(self.lib.lux_b200_step_host_sparse(
            ctypes.byref(self._desc), entries_host.data_ptr(), int(k), self.unit_actions.data_ptr(),
            self.tile_actions.data_ptr(), done_host.data_ptr() if done_host is not None else None,
            summary_host.data_ptr() if summary_host is not None else None,
            _lib.LUX_STEP_PREVALIDATE if prevalidate else 0, self._stream()), "lux_b200_step_host_sparse")

    def step_host_sparse_submit(self, entries_host, k, prevalidate=False):
        """First half of `step_host_sparse`: upload the list and queue scatter + step + done-flag pass; returns
        without waiting, so the caller can queue more work (pool resets, observations) behind it."""
        _lib.check(self.lib.lux_b200_step_host_sparse_submit(
            ctypes.byref(self._desc), entries_host.data_ptr(), int(k), self.unit_actions.data_ptr(),
            self.tile_actions.data_ptr(), _lib.LUX_STEP_PREVALIDATE if prevalidate else 0, self._stream()),
            "lux_b200_step_host_sparse_submit")

    def step_host_wait(self, done_host=None, summary_host=None):
        """Second half: block until the submitted turn's done flags and summary are in host memory."""
        _lib.check(self.lib.lux_b200_step_host_wait(
            ctypes.byref(self._desc), done_host.data_ptr() if done_host is not None else None,
            summary_host.data_ptr() if summary_host is not None else None, self._stream()), "lux_b200_step_host_wait")

    def pack_entries(self, unit_actions=None, tile_actions=None):
        """Device helper: the non-zero live entries of dense action tensors as an int32 [k, 2] entry list
        (the format step_host_sparse uploads).  Rows beyond the live prefixes are ignored."""
        ua = self.unit_actions if unit_actions is None else unit_actions
        ta = self.tile_actions if tile_actions is None else tile_actions
        hdr = self.hdr.view(torch.int16)
        nu = hdr[:, 6].to(torch.int64) & 0xFFFF          # n_units at byte 12
        nt = hdr[:, 8].to(torch.int64) & 0xFFFF          # n_tiles at byte 16
        live = self.hdr[:, 24] == 0
        um = (torch.arange(ua.shape[1], device=self.device)[None, :] < nu[:, None]) & (ua != 0) & live[:, None]
        tm = (torch.arange(ta.shape[1], device=self.device)[None, :] < nt[:, None]) & (ta != 0) & live[:, None]
        ui, ti = um.nonzero(), tm.nonzero()
        ukey = (ui[:, 0] << 10) | ui[:, 1]
        tkey = (1 << 31) | (ti[:, 0] << 10) | ti[:, 1]
        keys = torch.cat([ukey, tkey])
        vals = torch.cat([ua[um].to(torch.int64) & 0xFFFFFFFF, ta[tm].to(torch.int64)])
        ent = torch.stack([keys, vals], dim=1) & 0xFFFFFFFF
        return ((ent + 2 ** 31) % 2 ** 32 - 2 ** 31).to(torch.int32)      # the low 32 bits, as int32

    def gen_actions(self, seed, match_id_base=0, unit_actions=None, tile_actions=None, counters=None):
        """Fill the action tensors from the canonical seeded stream (harness utility)."""
        ua = self.unit_actions if unit_actions is None else unit_actions
        ta = self.tile_actions if tile_actions is None else tile_actions
        _lib.check(self.lib.lux_b200_gen_actions(ctypes.byref(self._desc), seed, match_id_base,
                                                 ua.data_ptr(), ta.data_ptr(),
                                                 counters.data_ptr() if counters is not None else None,
                                                 self._stream()), "lux_b200_gen_actions")

    def reset_done(self, pool, epoch, reset_count=None):
        """Replace finished matches by pool entries chosen by hash(match, epoch)."""
        _lib.check(self.lib.lux_b200_reset_done(ctypes.byref(self._desc), ctypes.byref(pool._desc),
                                                int(epoch) & 0xFFFFFFFF,
                                                reset_count.data_ptr() if reset_count is not None else None,
                                                self._stream()), "lux_b200_reset_done")

    def reset_done_dev(self, pool, epoch_dev, reset_count=None):
        """reset_done with the epoch in a device uint32 / int32 tensor (read when the kernel runs)."""
        _lib.check(self.lib.lux_b200_reset_done_dev(ctypes.byref(self._desc), ctypes.byref(pool._desc), epoch_dev.data_ptr(),
                                                    reset_count.data_ptr() if reset_count is not None else None,
                                                    self._stream()), "lux_b200_reset_done_dev")

    def reset_gen(self, pool, epoch, seed, match_id_base=0, reset_count=None, counters=None, unit_actions=None,
                  tile_actions=None):
        """reset_done(pool, epoch) followed by gen_actions(seed, match_id_base) in one kernel (harness loops)."""
        ua = self.unit_actions if unit_actions is None else unit_actions
        ta = self.tile_actions if tile_actions is None else tile_actions
        _lib.check(self.lib.lux_b200_reset_gen(ctypes.byref(self._desc), ctypes.byref(pool._desc), int(epoch) & 0xFFFFFFFF,
                                               reset_count.data_ptr() if reset_count is not None else None,
                                               seed, match_id_base, ua.data_ptr(), ta.data_ptr(),
                                               counters.data_ptr() if counters is not None else None, self._stream()),
                   "lux_b200_reset_gen")

    def observe(self, team_sel, e_cap=None):
        """Per-entity observation rows for team `team_sel[m]` of every match.

        Returns (obs float32 [n, e_cap, 85], ent int16 [n, e_cap], count int32 [n]); entity codes are
        unit slots, or 0x4000 | tile position for city tiles.  Rows beyond count[m] are unspecified.
        """
        e_cap = int(e_cap or (self.caps["units"] + self.caps["tiles"]))
        key = ("obs", e_cap)
        if getattr(self, "_obs_key", None) != key:
            self._obs = torch.empty((self.n, e_cap, 85), dtype=torch.float32, device=self.device)
            self._ent = torch.empty((self.n, e_cap), dtype=torch.int16, device=self.device)
            self._cnt = torch.zeros(self.n, dtype=torch.int32, device=self.device)
            self._obs_key = key
        if not torch.is_tensor(team_sel):
            team_sel = torch.full((self.n,), int(team_sel), dtype=torch.uint8, device=self.device)
        assert team_sel.dtype == torch.uint8 and team_sel.numel() == self.n and team_sel.is_cuda
        _lib.check(self.lib.lux_b200_observe(ctypes.byref(self._desc), team_sel.data_ptr(), self._obs.data_ptr(),
                                             self._ent.data_ptr(), self._cnt.data_ptr(), e_cap, None, self._stream()),
                   "lux_b200_observe")
        return self._obs, self._ent, self._cnt

    def observe_packed(self, team_sel, row_cap=None):
        """Observation rows without padding: returns (obs [R, 85], ent [R] int16 (-1 = unused row),
        count [n] int32, offsets [n+1] int64).  Match m owns rows offsets[m] .. offsets[m+1]: an upper
        bound of its team's units + tiles (from the headers), of which the first count[m] are entities.
        With `row_cap` the buffers have exactly that many rows and nothing is read back to the host (the call
        can be captured in a CUDA graph): offsets are clamped to row_cap, so matches that would not fit get
        no rows that turn -- `self.row_overflow` (device int64) counts the rows that were cut."""
        if not torch.is_tensor(team_sel):
            team_sel = torch.full((self.n,), int(team_sel), dtype=torch.uint8, device=self.device)
        if row_cap is not None:
            row_cap = int(row_cap)
            if getattr(self, "_fixed_rows", 0) != row_cap:
                self._fobs = torch.zeros((row_cap, 85), dtype=torch.float32, device=self.device)
                self._fent = torch.full((row_cap,), -1, dtype=torch.int16, device=self.device)
                self._fcnt = torch.zeros(self.n, dtype=torch.int32, device=self.device)
                self._foff = torch.zeros(self.n + 1, dtype=torch.int64, device=self.device)
                self.row_overflow = torch.zeros((), dtype=torch.int64, device=self.device)
                self._fixed_rows = row_cap
            _lib.check(self.lib.lux_b200_observe_offsets(ctypes.byref(self._desc), team_sel.data_ptr(), row_cap,
                                                         self._foff.data_ptr(), self.row_overflow.data_ptr(), self._stream()),
                       "lux_b200_observe_offsets")
            _lib.check(self.lib.lux_b200_observe(ctypes.byref(self._desc), team_sel.data_ptr(), self._fobs.data_ptr(),
                                                 self._fent.data_ptr(), self._fcnt.data_ptr(), 0, self._foff.data_ptr(),
                                                 self._stream()), "lux_b200_observe")
            return self._fobs, self._fent, self._fcnt, self._foff
        h16 = self.hdr.view(torch.int16)
        sel = team_sel.to(torch.int64)
        tu = torch.gather(h16[:, 10:12].to(torch.int64) & 0xFFFF, 1, sel[:, None])[:, 0]      # n_team_units
        tt = torch.gather(h16[:, 50:52].to(torch.int64) & 0xFFFF, 1, sel[:, None])[:, 0]      # n_team_tiles
        offsets = torch.zeros(self.n + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(tu + tt, 0, out=offsets[1:])
        rows = getattr(self, "_packed_rows", 0)
        need = int(offsets[-1].item())
        if need > rows:
            rows = max(need * 5 // 4, 1024)
            self._pobs = torch.empty((rows, 85), dtype=torch.float32, device=self.device)
            self._pent = torch.empty(rows, dtype=torch.int16, device=self.device)
            self._packed_rows = rows
        if getattr(self, "_cnt", None) is None or self._cnt.numel() != self.n:
            self._cnt = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.lux_b200_observe(ctypes.byref(self._desc), team_sel.data_ptr(), self._pobs.data_ptr(),
                                             self._pent.data_ptr(), self._cnt.data_ptr(), 0, offsets.data_ptr(),
                                             self._stream()), "lux_b200_observe")
        return self._pobs[:need], self._pent[:need], self._cnt, offsets

    def encode_actions(self, ent, count, codes, unit_actions=None, tile_actions=None, offsets=None):
        """AgentPolicy action codes (int32 [n, e_cap], or [R] with the packed `offsets`) -> action tensors."""
        ua = self.unit_actions if unit_actions is None else unit_actions
        ta = self.tile_actions if tile_actions is None else tile_actions
        assert codes.dtype == torch.int32 and codes.is_contiguous() and codes.shape == ent.shape
        _lib.check(self.lib.lux_b200_encode_actions(ctypes.byref(self._desc), ent.data_ptr(), count.data_ptr(),
                                                    codes.data_ptr(), ent.shape[1] if offsets is None else 0,
                                                    offsets.data_ptr() if offsets is not None else None,
                                                    ua.data_ptr(), ta.data_ptr(), self._stream()),
                   "lux_b200_encode_actions")

    def unit_word(self, m, slot, code):
        """The packed action word AgentPolicy's unit action `code` decodes to for the unit in `slot` of match
        `m` (lux_b200_encode_actions on a one-row entity list; used by the single-match adapters)."""
        if getattr(self, "_one", None) is None:
            z = lambda dt: torch.zeros((self.n, 1), dtype=dt, device=self.device)
            self._one = (z(torch.int16), z(torch.int32), torch.zeros(self.n, dtype=torch.int32, device=self.device),
                         torch.zeros_like(self.unit_actions), torch.zeros_like(self.tile_actions))
        ent, codes, cnt, ua, ta = self._one
        cnt.zero_()
        ent[m, 0], codes[m, 0], cnt[m] = int(slot), int(code), 1
        self.encode_actions(ent, cnt, codes, unit_actions=ua, tile_actions=ta)
        return int(ua[m, int(slot)].item()) & 0xFFFFFFFF

    def env_finish(self, pool, team_sel, reward_state, reward, done, turn_dev, reset_count=None):
        """End of an environment turn in one kernel (lux_b200_env_finish): reward [n] float64 and done [n] uint8 of
        the turn just played; finished matches are reset from `pool`, their reward state cleared and their
        learning team drawn (team_sel updated in place).  turn_dev: device int32 / uint32 counter."""
        _lib.check(self.lib.lux_b200_env_finish(ctypes.byref(self._desc), ctypes.byref(pool._desc), team_sel.data_ptr(),
                                                reward_state.data_ptr(), reward.data_ptr(), done.data_ptr(),
                                                turn_dev.data_ptr(), reset_count.data_ptr() if reset_count is not None else None,
                                                self._stream()), "lux_b200_env_finish")

    def reward(self, team_sel, reward_state, out=None):
        """Per-turn AgentPolicy reward (float64 [n]); reward_state is int32 [n, 4], updated in place."""
        if out is None:
            out = torch.empty(self.n, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.lux_b200_reward(ctypes.byref(self._desc), team_sel.data_ptr(), reward_state.data_ptr(),
                                            out.data_ptr(), self._stream()), "lux_b200_reward")
        return out

    def stats(self):
        """Device int64[16] episode statistics (see include/lux_b200.h lux_b200_stats)."""
        _lib.check(self.lib.lux_b200_stats(ctypes.byref(self._desc), self._stats.data_ptr(), self._stream()),
                   "lux_b200_stats")
        return self._stats
