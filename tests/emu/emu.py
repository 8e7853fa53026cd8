"""TEST INFRASTRUCTURE ONLY -- builds/loads the host emulation of the CUDA turn engine.

tests/emu/emu_step.cpp compiles luxpythonenvgym_b200/csrc/lux_step_core.cuh with g++ and a
32-fiber warp emulator, so the kernel's logic can be diffed against the oracle on CPU.
Nothing in the product imports this.
"""
import ctypes
import os
import subprocess

import numpy as np

from oracle.coracle import LuxBatchC, _p

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = None
GROUP = 32     # lane-group width the emulated step runs with (8 / 16 / 32); tests sweep it


def lib():
    global _LIB
    if _LIB is None:
        out = os.path.join(_HERE, "_build", "liblux_emu.so")
        csrc = os.path.join(_ROOT, "luxpythonenvgym_b200", "csrc")
        srcs = [os.path.join(_HERE, "emu_step.cpp"), os.path.join(_HERE, "warp_emu.h"),
                os.path.join(_ROOT, "include", "lux_b200.h")]
        srcs += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
        if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.run(["g++", "-O2", "-g", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas",
                            "-I", _HERE, "-o", out, srcs[0]], check=True)
        _LIB = ctypes.CDLL(out)
    return _LIB


def step(ms, unit_words, tile_bytes, flags=0, group=None):
    """One turn of one MatchState in place through the emulated kernel."""
    L = lib()
    hdr = ms.hdr.reshape(1)
    uw = np.zeros(len(ms.units), dtype=np.uint32)
    tb = np.zeros(len(ms.tiles), dtype=np.uint8)
    if unit_words is not None:
        uw[:] = unit_words[: len(uw)]
    if tile_bytes is not None:
        tb[:] = tile_bytes[: len(tb)]
    b = LuxBatchC(1, ms.size, len(ms.units), len(ms.cities), len(ms.tiles), len(ms.res),
                  hdr.ctypes.data, ms.cells.ctypes.data, ms.units.ctypes.data, ms.cities.ctypes.data,
                  ms.tiles.ctypes.data, ms.res.ctypes.data)
    rc = L.lux_emu_step_group(ctypes.byref(b), _p(uw), _p(tb), int(flags), int(group or GROUP))
    ms.hdr = hdr[0]
    if rc != 0:
        raise RuntimeError("lux_emu_step -> %d" % rc)
    return bool(ms.hdr["done"])


def step_many(states, unit_words, tile_bytes, flags=0, group=None):
    """One turn of several same-size MatchStates in place as ONE batch: with group < 32 they share
    warps (32 / group lane groups in lockstep), which is what the multi-match kernel runs."""
    L = lib()
    n, ms0 = len(states), states[0]
    arr = {k: np.ascontiguousarray(np.stack([getattr(ms, k).view(np.uint8).reshape(-1) for ms in states]))
           for k in ("cells", "units", "cities", "tiles", "res")}
    arr["hdr"] = np.ascontiguousarray(np.stack([ms.hdr.reshape(1).view(np.uint8).reshape(-1) for ms in states]))
    uw = np.zeros((n, len(ms0.units)), dtype=np.uint32)
    tb = np.zeros((n, len(ms0.tiles)), dtype=np.uint8)
    for i in range(n):
        uw[i] = unit_words[i][: uw.shape[1]]
        tb[i] = tile_bytes[i][: tb.shape[1]]
    b = LuxBatchC(n, ms0.size, len(ms0.units), len(ms0.cities), len(ms0.tiles), len(ms0.res),
                  *[arr[k].ctypes.data for k in ("hdr", "cells", "units", "cities", "tiles", "res")])
    rc = L.lux_emu_step_group(ctypes.byref(b), _p(uw), _p(tb), int(flags), int(group or GROUP))
    if rc != 0:
        raise RuntimeError("lux_emu_step_group -> %d" % rc)
    for i, ms in enumerate(states):
        ms.hdr = arr["hdr"][i].view(ms.hdr.dtype)[0].copy()
        for k in ("cells", "units", "cities", "tiles", "res"):
            dst = getattr(ms, k)
            dst[...] = arr[k][i].view(dst.dtype).reshape(dst.shape)


def need(ms):
    """(LuxNeed fields u, c, t, r, q, slice bytes) the kernel computes from this state's header."""
    L = lib()
    hdr = ms.hdr.reshape(1)
    out = np.zeros(5, dtype=np.int32)
    total = L.lux_emu_need(hdr.ctypes.data_as(ctypes.c_void_p), ms.size, len(ms.units), len(ms.cities), len(ms.tiles), len(ms.res), _p(out))
    return tuple(int(x) for x in out) + (int(total),)


def step_slice(ms, unit_words, tile_bytes, slice_bytes, flags=0):
    """One turn with a `slice_bytes` shared-memory slice: returns True when the turn did not fit (state untouched).
    A canary behind the slice catches any write past it."""
    L = lib()
    hdr = ms.hdr.reshape(1)
    uw = np.zeros(len(ms.units), dtype=np.uint32)
    tb = np.zeros(len(ms.tiles), dtype=np.uint8)
    uw[:] = unit_words[: len(uw)]
    tb[:] = tile_bytes[: len(tb)]
    b = LuxBatchC(1, ms.size, len(ms.units), len(ms.cities), len(ms.tiles), len(ms.res),
                  hdr.ctypes.data, ms.cells.ctypes.data, ms.units.ctypes.data, ms.cities.ctypes.data,
                  ms.tiles.ctypes.data, ms.res.ctypes.data)
    rc = L.lux_emu_step_slice(ctypes.byref(b), _p(uw), _p(tb), int(flags), int(slice_bytes))
    ms.hdr = hdr[0]
    if rc < 0:
        raise RuntimeError("lux_emu_step_slice -> %d" % rc)
    return rc > 0


def observe(ms, team, e_cap=640):
    """Observation rows of one MatchState through the emulated kernel."""
    L = lib()
    hdr = ms.hdr.reshape(1)
    b = LuxBatchC(1, ms.size, len(ms.units), len(ms.cities), len(ms.tiles), len(ms.res),
                  hdr.ctypes.data, ms.cells.ctypes.data, ms.units.ctypes.data, ms.cities.ctypes.data,
                  ms.tiles.ctypes.data, ms.res.ctypes.data)
    obs = np.zeros((e_cap, 85), dtype=np.float32)
    ent = np.zeros(e_cap, dtype=np.int16)
    cnt = np.zeros(1, dtype=np.int32)
    sel = np.array([team], dtype=np.uint8)
    rc = L.lux_emu_observe(ctypes.byref(b), _p(sel), _p(obs), _p(ent), _p(cnt), e_cap)
    if rc != 0:
        raise RuntimeError("lux_emu_observe -> %d" % rc)
    return obs[: cnt[0]].copy(), ent[: cnt[0]].copy()


def unit_word(ms, slot, code):
    """The packed action word AgentPolicy's unit action `code` decodes to for the unit in `slot`
    (csrc/lux_env_core.cuh: moves, smart transfers, build city, pillage)."""
    L = lib()
    hdr = ms.hdr.reshape(1)
    b = LuxBatchC(1, ms.size, len(ms.units), len(ms.cities), len(ms.tiles), len(ms.res),
                  hdr.ctypes.data, ms.cells.ctypes.data, ms.units.ctypes.data, ms.cities.ctypes.data,
                  ms.tiles.ctypes.data, ms.res.ctypes.data)
    L.lux_emu_unit_word.restype = ctypes.c_uint
    return int(L.lux_emu_unit_word(ctypes.byref(b), 0, int(slot), int(code)))
