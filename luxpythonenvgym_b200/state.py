"""Packed per-match state layout shared by the CUDA engine, the C-ABI and the host code.

One match is five flat little-endian sections; a batch is the same sections with a
leading match dimension (`[n_matches, ...]`, one torch uint8 tensor per section):

  hdr    LuxHeader            128 B   turn, id counters, counts, research, stats, flags
  cells  LuxCell[S*S]         8 B     resource / road / city-tile fields of every cell
  units  LuxUnit[U_cap]       16 B    team-A units in dict order, then team-B units
  cities LuxCity[C_cap]       16 B    cities in `Game.cities` dict order
  tiles  uint16[T_cap]        2 B     cell index (y*S+x) of every city tile in
                                      cities x city_cells order
  res    LuxRes[R_cap]        4 B     every resource cell in `GameMap.resources` order (static after
                                      reset): [9:0] cell index, [11:10] resource type, [31:16] a MIRROR
                                      of the cell's amount (kept equal to cells[cell].amount by every writer)

The field inventory follows the reference's state:
luxai2021/game/game.py:85-147 (state/stats/id counters), cell.py:20-33, unit.py:15-35,
city.py:15-66, game_map.py:53-58.  Cooldowns and road levels are multiples of 0.25 in
the reference and are stored as integer quarter units.  `include/lux_b200.h` declares the
same structs for C/CUDA; tests/test_abi.py checks the two agree.
"""
import numpy as np

HDR_BYTES = 128

# resource type codes (0 = none / depleted)
RES_NONE, RES_WOOD, RES_COAL, RES_URANIUM = 0, 1, 2, 3
RES_NAMES = {RES_WOOD: "wood", RES_COAL: "coal", RES_URANIUM: "uranium"}
RES_CODES = {v: k for k, v in RES_NAMES.items()}

# unit action word (uint32), one per unit slot
UA_NONE, UA_MOVE, UA_BCITY, UA_PILLAGE, UA_TRANSFER = 0, 1, 2, 3, 4
# move directions, bits [5:3] of a move word
DIR_CENTER, DIR_NORTH, DIR_EAST, DIR_SOUTH, DIR_WEST = 0, 1, 2, 3, 4
DIR_CHARS = "cnesw"
# city-tile action byte, one per tile in canonical order
TA_NONE, TA_BUILD_WORKER, TA_BUILD_CART, TA_RESEARCH = 0, 1, 2, 3

# per-match status bits (hdr.status)
ST_UNIT_OVERFLOW, ST_CITY_OVERFLOW, ST_TILE_OVERFLOW, ST_BAD_STATE = 1, 2, 4, 8

WIN_A, WIN_B, WIN_TIE, WIN_NONE = 0, 1, 2, 255

header_dtype = np.dtype([
    ("turn", "<i4"),
    ("unit_id_count", "<i4"),
    ("city_id_count", "<i4"),
    ("n_units", "<u2"),
    ("n_cities", "<u2"),
    ("n_tiles", "<u2"),
    ("n_res", "<u2"),
    ("n_team_units", "<u2", (2,)),
    ("done", "u1"),
    ("winner", "u1"),
    ("status", "u1"),
    ("size", "u1"),
    ("research", "<i4", (2,)),
    ("fuel_generated", "<i4", (2,)),
    ("collected", "<i4", (2, 3)),       # [team][wood, coal, uranium]
    ("workers_built", "<i4", (2,)),
    ("carts_built", "<i4", (2,)),
    ("roads_built_q", "<i4", (2,)),     # stats.roadsBuilt in quarter units
    ("city_tiles_built", "<i4", (2,)),  # only process_updates touches it (game.py:247)
    ("n_team_tiles", "<u2", (2,)),      # city tiles per team (derived; kept for rewards / win rule)
    ("pad", "u1", (24,)),
])
assert header_dtype.itemsize == HDR_BYTES

cell_dtype = np.dtype([
    ("amount", "<u2"),      # resource amount; 0 == no resource
    ("flags", "u1"),        # [1:0] resource type, [3:2] tile team+1 (0 none), [6:4] adjacent_city_tiles
    ("road_q", "u1"),       # raw Cell.road in quarter units
    ("tile_cd", "u1"),      # CityTile.cooldown (integer)
    ("city_slot", "u1"),    # index into cities[] when a tile
    ("key", "<u2"),         # index of the cell in its city's city_cells
])
assert cell_dtype.itemsize == 8

unit_dtype = np.dtype([
    ("id", "<i4"),          # N of "u_N"
    ("x", "u1"),
    ("y", "u1"),
    ("team_type", "u1"),    # bit0 team, bit1 cart
    ("cd_q", "u1"),         # cooldown in quarter units
    ("wood", "<u2"),
    ("coal", "<u2"),
    ("uranium", "<u2"),
    ("pad", "<u2"),
])
assert unit_dtype.itemsize == 16

city_dtype = np.dtype([
    ("id", "<i4"),          # N of "c_N"
    ("fuel", "<i4"),
    ("ntiles", "<u2"),
    ("adjsum", "<u2"),      # sum of adjacent_city_tiles over the city's tiles
    ("team", "u1"),
    ("pad", "u1", (3,)),
])
assert city_dtype.itemsize == 16


def default_caps(size):
    """Capacities (units, cities, tiles, resource cells) for a map side.

    Constraints of the C ABI: units % 4 == 0 and <= 252, cities <= 255, tiles % 16 == 0,
    res % 8 == 0 (16-byte granules of the bulk copies).  The largest populations in the
    reference's fixtures are 112 units / 133 tiles / 70 cities on 32x32.
    """
    cells = size * size

    def up(x, m):
        return (x + m - 1) // m * m

    units = up(min(252, max(64, cells // 2)), 4)
    tiles = up(min(384, max(64, cells)), 16)
    cities = min(255, max(32, cells // 4))
    # resource cells per generated map (20000 seeds per side): max 48 / 60 / 104 / 166 on 12 / 16 / 24 / 32
    res = {12: 64, 16: 80, 24: 128, 32: 192}.get(size, up(min(256, max(64, cells // 4)), 8))
    return dict(units=units, cities=cities, tiles=tiles, res=res)


def res_cell(entries):
    """Cell index of LuxRes entries (include/lux_b200.h)."""
    return np.asarray(entries, dtype=np.uint32) & 0x3FF


def res_type(entries):
    return (np.asarray(entries, dtype=np.uint32) >> 10) & 3


def res_amount(entries):
    return np.asarray(entries, dtype=np.uint32) >> 16


def res_make(cell, rtype, amount):
    return (np.asarray(cell, dtype=np.uint32) | (np.asarray(rtype, dtype=np.uint32) << 10) |
            (np.asarray(amount, dtype=np.uint32) << 16)).astype("<u4")


def encode_move(direction):
    return UA_MOVE | (direction << 3)


def encode_transfer(dest_id, res_code, amount):
    """res_code is RES_WOOD/RES_COAL/RES_URANIUM; amount is clamped to 2047 (a cart holds 2000)."""
    amount = max(0, min(int(amount), 2047))
    return UA_TRANSFER | ((res_code - 1) << 3) | (amount << 6) | (int(dest_id) << 17)


def decode_unit_action(word):
    word = int(word)
    kind = word & 7
    out = {"kind": kind}
    if kind == UA_MOVE:
        out["dir"] = (word >> 3) & 7
    elif kind == UA_TRANSFER:
        out["res"] = ((word >> 3) & 3) + 1
        out["amount"] = (word >> 6) & 2047
        out["dest"] = word >> 17
    return out


class MatchState:
    """One match's sections as numpy arrays (host side view used by tests, import/export)."""

    def __init__(self, size, caps=None):
        caps = dict(default_caps(size), **(caps or {}))
        self.size = size
        self.caps = caps
        self.hdr = np.zeros((), dtype=header_dtype)
        self.cells = np.zeros(size * size, dtype=cell_dtype)
        self.units = np.zeros(caps["units"], dtype=unit_dtype)
        self.cities = np.zeros(caps["cities"], dtype=city_dtype)
        self.tiles = np.zeros(caps["tiles"], dtype="<u2")
        self.res = np.zeros(caps["res"], dtype="<u4")
        self.hdr["size"] = size
        self.hdr["winner"] = WIN_NONE

    SECTIONS = ("hdr", "cells", "units", "cities", "tiles", "res")

    def live(self):
        """Dict of the canonical live state (what equality is defined over).

        `res` keeps only cells that still hold a resource (the reference drops depleted
        cells from GameMap.resources every turn, game.py:498-510, while the stored list is
        static), a depleted cell's stale resource type is masked, and hdr.n_res is ignored.
        """
        h = self.hdr.copy()
        h["n_res"] = 0
        cells = self.cells.copy()
        cells["flags"] = np.where(cells["amount"] == 0, cells["flags"] & 0xFC, cells["flags"])
        res = res_cell(self.res[: int(self.hdr["n_res"])])
        res = res[cells["amount"][res] > 0].astype("<u2")      # the list as the reference holds it: cell indices
        return {
            "hdr": h,
            "cells": cells,
            "units": self.units[: int(h["n_units"])],
            "cities": self.cities[: int(h["n_cities"])],
            "tiles": self.tiles[: int(h["n_tiles"])],
            "res": res,
        }

    def set_res_cells(self, cell_indices):
        """Fill the resource list from cell indices in GameMap.resources order; type and amount mirror come
        from the grid."""
        idx = np.asarray(cell_indices, dtype=np.int64).reshape(-1)
        if len(idx) > len(self.res):
            raise ValueError("more resource cells than the capacity %d" % len(self.res))
        self.res[:] = 0
        self.res[: len(idx)] = res_make(idx, self.cells["flags"][idx] & 3, self.cells["amount"][idx])
        self.hdr["n_res"] = len(idx)

    def mirror_errors(self):
        """Resource-list entries whose amount mirror (or, while the cell holds a resource, type) disagrees with
        the grid -- empty for every state this package or the oracle produces."""
        e = self.res[: int(self.hdr["n_res"])]
        cell = res_cell(e)
        amount = self.cells["amount"][cell]
        bad = (res_amount(e) != amount) | ((amount > 0) & (res_type(e) != (self.cells["flags"][cell] & 3)))
        return ["res[%d] (cell %d): mirror %d / type %d vs grid %d / type %d" % (
            k, cell[k], res_amount(e)[k], res_type(e)[k], amount[k], self.cells["flags"][cell[k]] & 3) for k in np.nonzero(bad)[0][:4]]

    def diff(self, other):
        """List of human-readable differences (empty == bit-identical live state)."""
        out = self.mirror_errors() + other.mirror_errors()
        a, b = self.live(), other.live()
        for name in self.SECTIONS:
            x, y = a[name], b[name]
            if x.shape != y.shape:
                out.append("%s: shape %s vs %s" % (name, x.shape, y.shape))
                continue
            if x.dtype.names:
                for f in x.dtype.names:
                    if f == "pad":
                        continue
                    xv, yv = np.atleast_1d(x[f]), np.atleast_1d(y[f])
                    if not np.array_equal(xv, yv):
                        idx = np.argwhere(xv != yv)[:4].tolist()
                        out.append("%s.%s differs at %s: %s vs %s" % (
                            name, f, idx, [xv[tuple(i)].tolist() for i in idx],
                            [yv[tuple(i)].tolist() for i in idx]))
            elif not np.array_equal(x, y):
                idx = np.argwhere(x != y)[:4].ravel().tolist()
                out.append("%s differs at %s: %s vs %s" % (name, idx, x[idx].tolist(), y[idx].tolist()))
        return out

    def digest(self):
        """64-bit digest of the canonical live state (golden fixtures store one per turn)."""
        import hashlib
        hh = hashlib.blake2b(digest_size=8)
        live = self.live()
        for name in self.SECTIONS:
            arr = live[name]
            if arr.dtype.names and "pad" in arr.dtype.names:
                arr = arr.copy()
                arr["pad"] = 0
            hh.update(np.ascontiguousarray(arr).tobytes())
        return int.from_bytes(hh.digest(), "little")

    @classmethod
    def from_arrays(cls, size, arrays, legacy_res=False):
        """Build from a dict of raw uint8/typed arrays keyed by section name.  `legacy_res`: the "res" array is
        the round-1 form (uint16 cell indices, as stored in tests/golden/*.npz); it is widened to LuxRes entries
        with type and amount taken from the grid."""
        caps = dict(units=arrays["units"].view(np.uint8).size // 16, cities=arrays["cities"].view(np.uint8).size // 16,
                    tiles=arrays["tiles"].view(np.uint8).size // 2, res=arrays["res"].view(np.uint8).size // (2 if legacy_res else 4))
        o = cls(size, caps)
        o.hdr = np.ascontiguousarray(arrays["hdr"]).view(np.uint8).reshape(-1).copy().view(header_dtype)[0]
        o.cells = np.ascontiguousarray(arrays["cells"]).view(np.uint8).reshape(-1).copy().view(cell_dtype)
        o.units = np.ascontiguousarray(arrays["units"]).view(np.uint8).reshape(-1).copy().view(unit_dtype)
        o.cities = np.ascontiguousarray(arrays["cities"]).view(np.uint8).reshape(-1).copy().view(city_dtype)
        o.tiles = np.ascontiguousarray(arrays["tiles"]).view(np.uint8).reshape(-1).copy().view("<u2")
        if legacy_res:
            o.set_res_cells(np.ascontiguousarray(arrays["res"]).view(np.uint8).reshape(-1).copy().view("<u2")[: int(o.hdr["n_res"])])
        else:
            o.res = np.ascontiguousarray(arrays["res"]).view(np.uint8).reshape(-1).copy().view("<u4")
        return o

    def copy(self):
        o = MatchState(self.size, self.caps)
        for name in self.SECTIONS:
            setattr(o, name, getattr(self, name).copy())
        return o

    def to_bytes(self):
        return {name: np.ascontiguousarray(getattr(self, name)).view(np.uint8).reshape(-1)
                for name in self.SECTIONS}
