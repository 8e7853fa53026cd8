"""Replay export from packed device states.

Writes the viewer JSON the reference's `Replay` produces (luxai2021/game/replay.py:27-80):
{seed, mapType, teamDetails, allCommands, version "3.1.0", results, width, height[, stateful]}.
Commands are the Kaggle wire texts of luxai2021/game/actions.py `to_message` (m / bw / bc / bcity / t / p / r),
reconstructed from the packed action tensors of one match; stateful snapshots follow
`Game.to_state_object` (game.py:1206-1271).  Host side only: export stays on the host like map generation.
"""
import json

from . import state as S


def commands_from_words(ms, unit_words, tile_bytes):
    """Packed action tensors of one match (state `ms` BEFORE the turn) -> [{"command", "agentID"}, ...]
    in the canonical action order (tiles in cities x city_cells order, then units in slot order)."""
    out = []
    size = ms.size
    for p in range(int(ms.hdr["n_tiles"])):
        code = int(tile_bytes[p])
        if code == S.TA_NONE:
            continue
        cell = int(ms.tiles[p])
        team = ((int(ms.cells["flags"][cell]) >> 2) & 3) - 1
        op = {S.TA_BUILD_WORKER: "bw", S.TA_BUILD_CART: "bc", S.TA_RESEARCH: "r"}[code]
        out.append({"command": "%s %d %d" % (op, cell % size, cell // size), "agentID": team})
    for i in range(int(ms.hdr["n_units"])):
        a = S.decode_unit_action(unit_words[i])
        k = a["kind"]
        if k == S.UA_NONE:
            continue
        team, uid = int(ms.units["team_type"][i]) & 1, "u_%d" % int(ms.units["id"][i])
        if k == S.UA_MOVE:
            cmd = "m %s %s" % (uid, S.DIR_CHARS[a["dir"]] if a["dir"] < 5 else "?")
        elif k == S.UA_BCITY:
            cmd = "bcity %s" % uid
        elif k == S.UA_PILLAGE:
            cmd = "p %s" % uid
        else:
            cmd = "t %s u_%d %s %d" % (uid, a["dest"], S.RES_NAMES[a["res"]], a["amount"])
        out.append({"command": cmd, "agentID": team})
    return out


def state_object(ms):
    """Game.to_state_object (game.py:1206-1271) of a packed MatchState."""
    size = ms.size
    h = ms.hdr
    cities, slots = {}, []
    for s in range(int(h["n_cities"])):
        c = ms.cities[s]
        cid = "c_%d" % int(c["id"])
        cities[cid] = {"id": cid, "fuel": int(c["fuel"]), "lightupkeep": 23 * int(c["ntiles"]) - 5 * int(c["adjsum"]),
                       "team": int(c["team"]), "cityCells": []}
        slots.append(cid)
    for p in range(int(h["n_tiles"])):
        cell = int(ms.tiles[p])
        c = ms.cells[cell]
        cities[slots[int(c["city_slot"])]]["cityCells"].append(
            {"x": cell % size, "y": cell // size, "cooldown": float(int(c["tile_cd"]))})
    teams = {}
    for t in (0, 1):
        rp = int(h["research"][t])
        teams[t] = {"researchPoints": rp, "units": {}, "researched": {"wood": True, "coal": rp >= 50, "uranium": rp >= 200}}
    for i in range(int(h["n_units"])):
        u = ms.units[i]
        teams[int(u["team_type"]) & 1]["units"]["u_%d" % int(u["id"])] = {
            "cargo": {"wood": int(u["wood"]), "uranium": int(u["uranium"]), "coal": int(u["coal"])},
            "cooldown": int(u["cd_q"]) / 4.0, "x": int(u["x"]), "y": int(u["y"]), "type": (int(u["team_type"]) >> 1) & 1}
    grid = []
    for y in range(size):
        row = []
        for x in range(size):
            c = ms.cells[y * size + x]
            d = {}
            road = 6 if (int(c["flags"]) & 0x0C) else int(c["road_q"]) / 4.0
            if road != 0:
                d["road"] = road
            if int(c["amount"]) > 0:
                d["type"], d["amount"] = S.RES_NAMES[int(c["flags"]) & 3], int(c["amount"])
            row.append(d)
        grid.append(row)
    return {"turn": int(h["turn"]), "globalCityIDCount": int(h["city_id_count"]),
            "globalunit_idCount": int(h["unit_id_count"]), "teamStates": teams, "map": grid, "cities": cities}


class Replay:
    """Accumulates one match's turns and writes the viewer JSON (reference: luxai2021/game/replay.py)."""

    def __init__(self, seed, size, stateful=False):
        self.data = {"seed": seed, "mapType": "random",
                     "teamDetails": [{"name": "Agent0", "tournamentID": ""}, {"name": "Agent1", "tournamentID": ""}],
                     "allCommands": [], "version": "3.1.0",
                     "results": {"ranks": [{"rank": 1, "agentID": 0}, {"rank": 2, "agentID": 1}], "replayFile": ""},
                     "width": size, "height": size}
        self.stateful = stateful
        if stateful:
            self.data["stateful"] = []

    def clear(self):
        """Forget the turns logged so far (Game.reset, reference replay.py:27-41)."""
        self.data["allCommands"] = []
        self.data["results"]["ranks"] = [{"rank": 1, "agentID": 0}, {"rank": 2, "agentID": 1}]
        if self.stateful:
            self.data["stateful"] = []

    def add_state(self, state_after):
        if self.stateful:
            self.data["stateful"].append(state_object(state_after))

    def add_turn(self, state_before, unit_words, tile_bytes, state_after=None):
        self.data["allCommands"].append(commands_from_words(state_before, unit_words, tile_bytes))
        if self.stateful and state_after is not None:
            self.data["stateful"].append(state_object(state_after))

    def set_result(self, winner):
        if winner == S.WIN_B:
            self.data["results"]["ranks"] = [{"rank": 1, "agentID": 1}, {"rank": 2, "agentID": 0}]

    def write(self, path):
        with open(path, "w") as f:
            json.dump(self.data, f)
        return path
