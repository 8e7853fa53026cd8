"""TEST INFRASTRUCTURE ONLY -- canonical seeded action stream, Python statement.

State-independent counter hash per entity: the same function is
restated in C (oracle/lux_oracle.c lux_oracle_gen_actions) and on the device
(lux_b200_gen_actions); tests check the three agree.  Works on a packed MatchState, so it
can drive the reference engine (through oracle/flatten.words_to_actions), the C oracle and
the CUDA engine with identical action tensors.
"""
import numpy as np

from luxpythonenvgym_b200 import state as S

M64 = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15


def mix64(z):
    z &= M64
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return z


def hash64(seed, match, turn, kind, key):
    x = seed & M64
    for v in (match, turn, kind, key):
        x = mix64(x + GOLD * (v + 1))
    return x


def gen_actions(ms, seed, match_id):
    h = ms.hdr
    uw = np.zeros(len(ms.units), dtype=np.uint32)
    tb = np.zeros(len(ms.tiles), dtype=np.uint8)
    if h["done"]:
        return uw, tb
    turn = int(h["turn"])
    nu, nt = int(h["n_units"]), int(h["n_tiles"])
    U = ms.units
    for i in range(nu):
        r = hash64(seed, match_id, turn, 0, int(U["id"][i]))
        can_act = U["cd_q"][i] < 4
        if not can_act and ((r >> 32) & 15) != 0:
            continue
        p = r % 100
        w = 0
        ucell = int(U["y"][i]) * ms.size + int(U["x"][i])
        on_tile = ((int(ms.cells["flags"][ucell]) >> 2) & 3) != 0
        cargo = int(U["wood"][i]) + int(U["coal"][i]) + int(U["uranium"][i])
        night = (turn % 40) >= 30
        can_build = (not (int(U["team_type"][i]) & 2)) and cargo >= 100 and not on_tile \
            and int(ms.cells["amount"][ucell]) == 0
        if can_build and ((r >> 40) & 3) != 0:
            w = S.UA_BCITY
        elif night and on_tile and ((r >> 42) & 7) != 0:
            w = 0
        elif p < 40:
            w = 0
        elif p < 80:
            w = S.encode_move(1 + ((r >> 8) & 3))
        elif p < 85:
            w = S.UA_BCITY
        elif p < 88:
            w = S.UA_PILLAGE
        else:
            best = -1
            for j in range(nu):
                if j == i or ((int(U["team_type"][j]) ^ int(U["team_type"][i])) & 1):
                    continue
                d = abs(int(U["x"][j]) - int(U["x"][i])) + abs(int(U["y"][j]) - int(U["y"][i]))
                if d <= 1 and (best < 0 or U["id"][j] < U["id"][best]):
                    best = j
            if best >= 0:
                w = S.encode_transfer(int(U["id"][best]), 1 + (r >> 8) % 3, 1 + (r >> 16) % 100)
        uw[i] = w
    for p in range(nt):
        cell = int(ms.tiles[p])
        r = hash64(seed, match_id, turn, 1, cell)
        can_act = ms.cells["tile_cd"][cell] < 1
        if not can_act and ((r >> 32) & 15) != 0:
            continue
        q = r % 100
        tb[p] = S.TA_BUILD_WORKER if q < 45 else S.TA_BUILD_CART if q < 50 else S.TA_RESEARCH if q < 90 else S.TA_NONE
    return uw, tb
