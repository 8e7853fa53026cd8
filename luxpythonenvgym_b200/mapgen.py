"""Seeded host-side map generation -> packed initial MatchState.

Map generation stays on the host (BASELINE.json north_star).  This module produces, for a
seed, exactly the initial state the reference builds in `Game.reset()` ->
`GameMap.generate_map` (luxai2021/game/game_map.py:60-194) so that seeded matches are
comparable turn by turn; tests/test_mapgen.py pins it against the 100 golden maps of
luxai2021/tests/testmaps.json (tests/golden/maps.npz).

Pipeline (game_map.py): size draw :86 -> symmetry :111 -> per resource type a random fill +
2 rounds of a life-like automaton :313-366 -> 10 rounds of like-attracts / unlike-repels
drift :368-438 -> 5 % perturbation :283-301 -> mirror :303-309 -> total-amount validation
with retry :196-214 -> spawn cell :143-147 -> up to six 800-wood cells next to the spawns
:158-192.  All arithmetic is IEEE double in the reference's evaluation order.

The generator is seedrandom's ARC4 stream seeded with the string "gen_<seed>"
(luxai2021/env/rng/rng.js:4, seedrandom.js:1).
"""
import math

import numpy as np

from . import state as S

MAP_SIZES = (12, 16, 24, 32)
# neighbour offsets in the reference's MOVE_DELTAS order (game_map.py:23-32)
DELTAS = ((0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1), (1, 0), (1, 1))
HORIZONTAL, VERTICAL = 0, 1
WOOD, COAL, URANIUM = S.RES_WOOD, S.RES_COAL, S.RES_URANIUM


class Arc4Random:
    """seedrandom(seed_string)() : RC4 keystream -> doubles in [0, 1)."""

    def __init__(self, seed_string):
        key = [ord(ch) & 255 for ch in seed_string] or [0]
        perm = list(range(256))
        j = 0
        for i in range(256):
            j = (j + perm[i] + key[i % len(key)]) & 255
            perm[i], perm[j] = perm[j], perm[i]
        self._perm, self._i, self._j = perm, 0, 0
        self._take(256)  # seedrandom discards the first 256 bytes

    def _take(self, nbytes):
        perm, i, j, acc = self._perm, self._i, self._j, 0
        for _ in range(nbytes):
            i = (i + 1) & 255
            j = (j + perm[i]) & 255
            perm[i], perm[j] = perm[j], perm[i]
            acc = (acc << 8) | perm[(perm[i] + perm[j]) & 255]
        self._i, self._j = i, j
        return acc

    def random(self):
        num = float(self._take(6))      # 48 bits
        den = 281474976710656.0         # 256**6
        extra = 0
        while num < 4503599627370496.0:  # < 2**52: pull in another byte
            num = (num + extra) * 256.0
            den *= 256.0
            extra = self._take(1)
        while num >= 9007199254740992.0:  # >= 2**53: shift back
            num /= 2.0
            den /= 2.0
            extra >>= 1
        return (num + extra) / den


class _Res:
    """A resource deposit while the map is being laid out (moved around by reference)."""
    __slots__ = ("kind", "amt", "fx", "fy")

    def __init__(self, kind, amt):
        self.kind, self.amt, self.fx, self.fy = kind, amt, 0.0, 0.0


def _sgn(v):
    return 1 if v > 0 else (-1 if v < 0 else 0)


def _automaton_layer(rnd, density, spread, w, h, death_limit, birth_limit):
    p = density - spread / 2 + spread * rnd()
    grid = [[1 if rnd() < p else 0 for _ in range(w)] for _ in range(h)]
    for _ in range(2):
        # in-place sweep over the interior, exactly as the reference updates while scanning
        for i in range(1, h - 1):
            row = grid[i]
            for j in range(1, w - 1):
                alive = 0
                for dx, dy in DELTAS:
                    alive += grid[i + dy][j + dx] == 1
                if row[j] == 1:
                    row[j] = 0 if alive < death_limit else 1
                else:
                    row[j] = 1 if alive > birth_limit else 0
    return grid


def _drift(field):
    h, w = len(field), len(field[0])
    for y in range(h):
        for x in range(w):
            r = field[y][x]
            if r is None:
                continue
            fx = fy = 0.0
            for yy in range(y - 5, y + 5):
                if yy < 0 or yy >= h:
                    continue
                for xx in range(x - 5, x + 5):
                    if xx < 0 or xx >= w:
                        continue
                    o = field[yy][xx]
                    if o is None:
                        continue
                    dx, dy = x - xx, y - yy
                    md = abs(dx) + abs(dy)
                    if o.kind != r.kind:
                        if dx:
                            fx += math.pow(dx / md, 2) * _sgn(dx)
                        if dy:
                            fy += math.pow(dy / md, 2) * _sgn(dy)
                    else:
                        if dx:
                            fx -= math.pow(dx / md, 2) * _sgn(dx)
                        if dy:
                            fy -= math.pow(dy / md, 2) * _sgn(dy)
            r.fx, r.fy = fx, fy
    out = [[None] * w for _ in range(h)]
    for y in range(h):
        for x in range(w):
            r = field[y][x]
            if r is None:
                continue
            nx = min(max(x + _sgn(r.fx), 0), w - 1)
            ny = min(max(y + _sgn(r.fy), 0), h - 1)
            if out[ny][nx] is None:
                out[ny][nx] = r
            else:
                out[y][x] = r
    return out


def _wood_amt(rnd):
    return min(300 + math.floor(rnd() * 100), 500)


def _coal_amt(rnd):
    return 350 + math.floor(rnd() * 75)


def _uranium_amt(rnd):
    return 300 + math.floor(rnd() * 50)


def _layout(rnd, symmetry, w, h, hw, hh):
    field = [[None] * w for _ in range(h)]
    for kind, density, spread, dl, bl, amount in (
            (WOOD, 0.21, 0.01, 2, 4, _wood_amt),
            (COAL, 0.11, 0.02, 2, 4, _coal_amt),
            (URANIUM, 0.055, 0.04, 1, 6, _uranium_amt)):
        layer = _automaton_layer(rnd, density, spread, hw, hh, dl, bl)
        for y in range(hh):
            for x in range(hw):
                if layer[y][x] == 1:
                    field[y][x] = _Res(kind, amount(rnd))
    for _ in range(10):
        field = _drift(field)
    for y in range(hh):
        for x in range(hw):
            r = field[y][x]
            if r is None:
                continue
            for dx, dy in DELTAS:
                nx, ny = x + dx, y + dy
                # the reference compares nx with the half HEIGHT and ny with the half WIDTH
                if nx < 0 or ny < 0 or nx >= hh or ny >= hw:
                    continue
                if rnd() < 0.05:
                    amt = _uranium_amt(rnd)
                    if r.kind == COAL:
                        amt = _coal_amt(rnd)
                    if r.kind == WOOD:
                        amt = _wood_amt(rnd)
                    field[ny][nx] = _Res(r.kind, amt)
    for y in range(hh):
        for x in range(hw):
            if symmetry == VERTICAL:
                field[y][w - x - 1] = field[y][x]
            else:
                field[h - y - 1][x] = field[y][x]
    return field


def _enough(field):
    tot = {WOOD: 0, COAL: 0, URANIUM: 0}
    for row in field:
        for r in row:
            if r is not None:
                tot[r.kind] += r.amt
    return tot[WOOD] >= 2000 and tot[COAL] >= 1500 and tot[URANIUM] >= 300


def generate(seed, size=None, caps=None):
    """Initial MatchState for `seed`; `size` forces width=height (configs["width"/"height"])."""
    rnd = Arc4Random("gen_%d" % int(seed)).random
    drawn = MAP_SIZES[math.floor(rnd() * len(MAP_SIZES))]
    w = h = int(size) if size is not None else drawn
    hw, hh = w, h
    if rnd() < 0.5:
        symmetry = VERTICAL
        hw = w // 2
    else:
        symmetry = HORIZONTAL
        hh = h // 2
    field = _layout(rnd, symmetry, w, h, hw, hh)
    while not _enough(field):
        field = _layout(rnd, symmetry, w, h, hw, hh)

    ms = S.MatchState(w, caps)
    res_list = []

    def add_resource(x, y, kind, amt):
        c = ms.cells[y * w + x]
        c["amount"] = amt
        c["flags"] = (int(c["flags"]) & 0xFC) | kind
        res_list.append(y * w + x)

    for y in range(h):
        for x in range(w):
            r = field[y][x]
            if r is not None:
                add_resource(x, y, r.kind, r.amt)

    sx = math.floor(rnd() * (hw - 1)) + 1
    sy = math.floor(rnd() * (hh - 1)) + 1
    while ms.cells["amount"][sy * w + sx] > 0:
        sx = math.floor(rnd() * (hw - 1)) + 1
        sy = math.floor(rnd() * (hh - 1)) + 1
    spawns = [(sx, sy), (sx, h - sy - 1) if symmetry == HORIZONTAL else (w - sx - 1, sy)]
    for team, (x, y) in enumerate(spawns):
        u = ms.units[team]
        u["id"], u["x"], u["y"], u["team_type"] = team + 1, x, y, team
        cty = ms.cities[team]
        cty["id"], cty["team"], cty["ntiles"] = team + 1, team, 1
        c = ms.cells[y * w + x]
        c["flags"] = int(c["flags"]) | ((team + 1) << 2)
        c["city_slot"] = team
        ms.tiles[team] = y * w + x

    first = math.floor(rnd() * len(DELTAS))
    placed = 0
    for k in range(7):
        dx, dy = DELTAS[(first + k) % len(DELTAS)]
        nx, ny = sx + dx, sy + dy
        nx2, ny2 = (nx, h - ny - 1) if symmetry == HORIZONTAL else (w - nx - 1, ny)
        if not (0 <= nx < w and 0 <= ny < h and 0 <= nx2 < w and 0 <= ny2 < h):
            continue
        for px, py in ((nx, ny), (nx2, ny2)):
            c = ms.cells[py * w + px]
            if c["amount"] == 0 and (int(c["flags"]) & 0x0C) == 0:
                placed += 1
                add_resource(px, py, WOOD, 800)
        if placed == 6:
            break

    hd = ms.hdr
    hd["unit_id_count"] = 2
    hd["city_id_count"] = 2
    hd["n_units"], hd["n_cities"], hd["n_tiles"] = 2, 2, 2
    hd["n_team_units"] = (1, 1)
    hd["n_team_tiles"] = (1, 1)
    hd["workers_built"] = (1, 1)
    assert len(res_list) <= len(ms.res), "resource capacity too small for this map"
    ms.set_res_cells(res_list)
    return ms


def natural_size(seed):
    """The map side the generator draws for this seed (first rng value, game_map.py:86)."""
    return MAP_SIZES[math.floor(Arc4Random("gen_%d" % int(seed)).random() * len(MAP_SIZES))]


# ---------------------------------------------------------------------------------------------------
# native generator (luxpythonenvgym_b200/csrc/lux_mapgen.cpp): same maps, ~100x faster, multithreaded
_native = None


def _native_lib():
    global _native
    if _native is None:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lux_mapgen.so")
        if not os.path.exists(path):
            from . import build
            build.build_mapgen()
        _native = ctypes.CDLL(path)
        _native.lux_mapgen_natural_size.argtypes = [ctypes.c_int64]
    return _native


def generate_batch(seeds, size, caps=None, threads=0):
    """Packed initial states for many seeds at once -> dict of host uint8 arrays [n, bytes] per section
    (feed to BatchedGame.load_arrays).  `size` forces the map side (all maps of a batch share it)."""
    import ctypes
    from ._lib import LuxBatch
    lib = _native_lib()
    caps = dict(S.default_caps(size), **(caps or {}))
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    n = len(seeds)
    arr = {name: np.zeros((n, width), dtype=np.uint8) for name, width in (
        ("hdr", S.HDR_BYTES), ("cells", size * size * 8), ("units", caps["units"] * 16),
        ("cities", caps["cities"] * 16), ("tiles", caps["tiles"] * 2), ("res", caps["res"] * 4))}
    desc = LuxBatch(n, size, caps["units"], caps["cities"], caps["tiles"], caps["res"],
                    *[arr[k].ctypes.data for k in ("hdr", "cells", "units", "cities", "tiles", "res")])
    rc = lib.lux_mapgen_generate_batch(ctypes.byref(desc), seeds.ctypes.data_as(ctypes.c_void_p), int(threads))
    if rc != 0:
        raise RuntimeError("lux_mapgen_generate_batch -> %d" % rc)
    return arr


def generate_native(seed, size=None, caps=None):
    """One map through the native generator, as a MatchState (size None = the seed's natural size)."""
    if size is None:
        size = int(_native_lib().lux_mapgen_natural_size(int(seed)))
    caps = dict(S.default_caps(size), **(caps or {}))
    arr = generate_batch([seed], size, caps, threads=1)
    return S.MatchState.from_arrays(size, {k: v[0] for k, v in arr.items()})
