"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/lux_oracle.c (the CPU checker)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return os.path.join(_HERE, "_build", "liblux_oracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liblux_oracle.so")
        src = os.path.join(_HERE, "lux_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.lux_oracle_hash.restype = ctypes.c_uint64
        _LIB.lux_oracle_hash.argtypes = [ctypes.c_uint64] * 5
    return _LIB


def _p(arr):
    return arr.ctypes.data_as(ctypes.c_void_p) if arr is not None else None


def step(ms, unit_words=None, tile_bytes=None):
    """One turn of one MatchState in place (lux_oracle_step)."""
    L = lib()
    hdr = ms.hdr.reshape(1)
    if unit_words is not None:
        unit_words = np.ascontiguousarray(unit_words, dtype=np.uint32)
        assert len(unit_words) >= len(ms.units)
    if tile_bytes is not None:
        tile_bytes = np.ascontiguousarray(tile_bytes, dtype=np.uint8)
        assert len(tile_bytes) >= len(ms.tiles)
    rc = L.lux_oracle_step(_p(hdr), _p(ms.cells), _p(ms.units), _p(ms.cities), _p(ms.tiles), _p(ms.res),
                           _p(unit_words), _p(tile_bytes), len(ms.units), len(ms.cities), len(ms.tiles))
    ms.hdr = hdr[0]
    if rc != 0:
        raise RuntimeError("lux_oracle_step -> %d" % rc)
    return bool(ms.hdr["done"])


def gen_actions(ms, seed, match_id):
    L = lib()
    uw = np.zeros(len(ms.units), dtype=np.uint32)
    tb = np.zeros(len(ms.tiles), dtype=np.uint8)
    hdr = ms.hdr.reshape(1)
    rc = L.lux_oracle_gen_actions(_p(hdr), _p(ms.cells), _p(ms.units), _p(ms.tiles),
                                  ctypes.c_uint64(seed), ctypes.c_uint64(match_id), _p(uw), _p(tb),
                                  len(ms.units), len(ms.tiles))
    if rc != 0:
        raise RuntimeError("lux_oracle_gen_actions -> %d" % rc)
    return uw, tb


class LuxBatchC(ctypes.Structure):
    _fields_ = [("n_matches", ctypes.c_int32), ("size", ctypes.c_int32),
                ("u_cap", ctypes.c_int32), ("c_cap", ctypes.c_int32),
                ("t_cap", ctypes.c_int32), ("r_cap", ctypes.c_int32),
                ("hdr", ctypes.c_void_p), ("cells", ctypes.c_void_p), ("units", ctypes.c_void_p),
                ("cities", ctypes.c_void_p), ("tiles", ctypes.c_void_p), ("res", ctypes.c_void_p)]


def batch_desc(size, caps, arrays, n):
    """arrays: dict of C-contiguous numpy uint8 arrays hdr/cells/units/cities/tiles/res with leading dim n."""
    b = LuxBatchC(n, size, caps["units"], caps["cities"], caps["tiles"], caps["res"],
                  *[arrays[k].ctypes.data for k in ("hdr", "cells", "units", "cities", "tiles", "res")])
    return b


def step_batch(desc, unit_actions, tile_actions, first=0, count=None):
    L = lib()
    count = desc.n_matches if count is None else count
    rc = L.lux_oracle_step_batch(ctypes.byref(desc), _p(unit_actions), _p(tile_actions), first, count)
    if rc != 0:
        raise RuntimeError("lux_oracle_step_batch -> %d" % rc)


def gen_actions_batch(desc, seed, match_id_base, unit_actions, tile_actions, first=0, count=None):
    L = lib()
    count = desc.n_matches if count is None else count
    rc = L.lux_oracle_gen_actions_batch(ctypes.byref(desc), ctypes.c_uint64(seed), ctypes.c_uint64(match_id_base),
                                        _p(unit_actions), _p(tile_actions), first, count)
    if rc != 0:
        raise RuntimeError("lux_oracle_gen_actions_batch -> %d" % rc)


def run_batch(desc, pool_desc, seed, match_id_base, unit_actions, tile_actions, first, count, turns, epoch0=0):
    """lux_oracle_run_batch: (step_seconds, live_match_turns) for matches [first, first+count)."""
    L = lib()
    secs = ctypes.c_double(0.0)
    live = ctypes.c_longlong(0)
    rc = L.lux_oracle_run_batch(ctypes.byref(desc), ctypes.byref(pool_desc) if pool_desc is not None else None,
                                ctypes.c_uint64(seed), ctypes.c_uint64(match_id_base), _p(unit_actions),
                                _p(tile_actions), first, count, turns, ctypes.c_uint32(epoch0),
                                ctypes.byref(secs), ctypes.byref(live))
    if rc != 0:
        raise RuntimeError("lux_oracle_run_batch -> %d" % rc)
    return secs.value, live.value


def observe(ms, team, e_cap=640):
    """lux_oracle_observe for one MatchState -> (obs float32 [E,85], ent codes int16 [E])."""
    L = lib()
    obs = np.zeros((e_cap, 85), dtype=np.float32)
    ent = np.zeros(e_cap, dtype=np.int16)
    cnt = ctypes.c_int32(0)
    hdr = ms.hdr.reshape(1)
    rc = L.lux_oracle_observe(_p(hdr), _p(ms.cells), _p(ms.units), _p(ms.cities), _p(ms.tiles), _p(ms.res),
                              int(team), _p(obs), _p(ent), ctypes.byref(cnt), e_cap)
    if rc != 0:
        raise RuntimeError("lux_oracle_observe -> %d" % rc)
    return obs[: cnt.value].copy(), ent[: cnt.value].copy()


def observe_batch(desc, team_sel, e_cap):
    L = lib()
    n = desc.n_matches
    obs = np.zeros((n, e_cap, 85), dtype=np.float32)
    ent = np.zeros((n, e_cap), dtype=np.int16)
    cnt = np.zeros(n, dtype=np.int32)
    team_sel = np.ascontiguousarray(team_sel, dtype=np.uint8)
    rc = L.lux_oracle_observe_batch(ctypes.byref(desc), _p(team_sel), _p(obs), _p(ent), _p(cnt), e_cap)
    if rc != 0:
        raise RuntimeError("lux_oracle_observe_batch -> %d" % rc)
    return obs, ent, cnt


def prevalidate(ms, unit_words, tile_bytes):
    """lux_oracle_prevalidate in place on copies; returns (unit_words, tile_bytes)."""
    L = lib()
    uw = np.ascontiguousarray(unit_words, dtype=np.uint32).copy()
    tb = np.ascontiguousarray(tile_bytes, dtype=np.uint8).copy()
    hdr = ms.hdr.reshape(1)
    rc = L.lux_oracle_prevalidate(_p(hdr), _p(ms.cells), _p(ms.units), _p(ms.tiles), _p(uw), _p(tb))
    if rc != 0:
        raise RuntimeError("lux_oracle_prevalidate -> %d" % rc)
    return uw, tb


def encode_actions(ms, ent, codes):
    L = lib()
    uw = np.zeros(len(ms.units), dtype=np.uint32)
    tb = np.zeros(len(ms.tiles), dtype=np.uint8)
    ent = np.ascontiguousarray(ent, dtype=np.int16)
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    hdr = ms.hdr.reshape(1)
    rc = L.lux_oracle_encode_actions(_p(hdr), _p(ms.units), _p(ent), len(ent), _p(codes), _p(uw), _p(tb),
                                     len(ms.units), len(ms.tiles))
    if rc != 0:
        raise RuntimeError("lux_oracle_encode_actions -> %d" % rc)
    return uw, tb


def reward(ms, team, last):
    """lux_oracle_reward; `last` is an int32[3] array updated in place."""
    L = lib()
    r = ctypes.c_double(0.0)
    hdr = ms.hdr.reshape(1)
    rc = L.lux_oracle_reward(_p(hdr), int(team), _p(last), ctypes.byref(r))
    if rc != 0:
        raise RuntimeError("lux_oracle_reward -> %d" % rc)
    return r.value
