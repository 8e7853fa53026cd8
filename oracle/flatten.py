"""TEST INFRASTRUCTURE ONLY -- reference `Game` object -> packed MatchState (and actions back).

`flatten(game)` walks the reference object model exactly in the orders the reference
itself iterates: team unit dicts (game.py:477-485), `Game.cities`
x `City.city_cells` (game.py:468-475), `GameMap.resources` (game_map.py:440-453), and
writes the canonical packed layout of luxpythonenvgym_b200/state.py.

`words_to_actions(game, unit_words, tile_bytes)` turns the packed action tensors of one
match into the reference's Action objects in the canonical list order (tiles in
cities x city_cells order, then units in slot order), i.e. the list that the device step
is defined to be equivalent to (luxai2021/game/game.py:390 `run_turn_with_actions`).
"""
import numpy as np

from luxpythonenvgym_b200 import state as S


def _q(value):
    q = value * 4
    qi = int(round(q))
    assert abs(q - qi) < 1e-9, "not a multiple of 0.25: %r" % (value,)
    return qi


def flatten(game, caps=None):
    size = game.map.width
    assert game.map.height == size, "square maps only"
    ms = S.MatchState(size, caps)
    h = ms.hdr
    h["turn"] = game.state["turn"]
    h["unit_id_count"] = game.global_unit_id_count
    h["city_id_count"] = game.global_city_id_count

    # cities + tiles, Game.cities dict order x city_cells order
    slot_of = {}
    ntile = 0
    for slot, city in enumerate(game.cities.values()):
        slot_of[city.id] = slot
        c = ms.cities[slot]
        c["id"] = int(city.id.split("_")[1])
        c["fuel"] = int(city.fuel)
        assert float(city.fuel) == int(city.fuel)
        c["team"] = city.team
        c["ntiles"] = len(city.city_cells)
        adjsum = 0
        for key, cell in enumerate(city.city_cells):
            ct = cell.city_tile
            assert ct is not None and ct.city_id == city.id
            idx = cell.pos.y * size + cell.pos.x
            ms.tiles[ntile] = idx
            ntile += 1
            cc = ms.cells[idx]
            assert float(ct.cooldown) == int(ct.cooldown)
            cc["tile_cd"] = int(ct.cooldown)
            cc["city_slot"] = slot
            cc["key"] = key
            cc["flags"] |= ((ct.team + 1) << 2) | (ct.adjacent_city_tiles << 4)
            adjsum += ct.adjacent_city_tiles
        c["adjsum"] = adjsum
    h["n_cities"] = len(game.cities)
    h["n_tiles"] = ntile
    for city in game.cities.values():
        h["n_team_tiles"][city.team] += len(city.city_cells)

    # cells: resources and roads
    for y in range(size):
        for x in range(size):
            cell = game.map.get_cell(x, y)
            cc = ms.cells[y * size + x]
            cc["road_q"] = _q(cell.road)
            if cell.resource is not None and cell.resource.amount > 0:
                cc["amount"] = cell.resource.amount
                cc["flags"] |= S.RES_CODES[cell.resource.type]
    ms.set_res_cells([cell.pos.y * size + cell.pos.x for cell in game.map.resources if cell.resource.amount > 0])

    # units: team A dict order, then team B
    n = 0
    for team in (0, 1):
        units = game.state["teamStates"][team]["units"]
        h["n_team_units"][team] = len(units)
        for unit in units.values():
            u = ms.units[n]
            n += 1
            u["id"] = int(unit.id.split("_")[1])
            u["x"], u["y"] = unit.pos.x, unit.pos.y
            u["team_type"] = unit.team | (unit.type << 1)
            u["cd_q"] = _q(unit.cooldown)
            u["wood"], u["coal"], u["uranium"] = unit.cargo["wood"], unit.cargo["coal"], unit.cargo["uranium"]
        h["research"][team] = game.state["teamStates"][team]["researchPoints"]
        st = game.stats["teamStats"][team]
        h["fuel_generated"][team] = st["fuelGenerated"]
        for k, name in enumerate(("wood", "coal", "uranium")):
            h["collected"][team][k] = st["resourcesCollected"][name]
        h["workers_built"][team] = st["workersBuilt"]
        h["carts_built"][team] = st["cartsBuilt"]
        h["roads_built_q"][team] = _q(st["roadsBuilt"])
        h["city_tiles_built"][team] = st["cityTilesBuilt"]
    h["n_units"] = n
    return ms


def winner_code(game):
    """get_winning_team (game.py:594-634) with the host-RNG coin flip reported as a tie."""
    tiles = [0, 0]
    for city in game.cities.values():
        tiles[city.team] += len(city.city_cells)
    if tiles[0] != tiles[1]:
        return S.WIN_A if tiles[0] > tiles[1] else S.WIN_B
    units = [len(game.state["teamStates"][t]["units"]) for t in (0, 1)]
    if units[0] != units[1]:
        return S.WIN_A if units[0] > units[1] else S.WIN_B
    fuel = [game.stats["teamStats"][t]["fuelGenerated"] for t in (0, 1)]
    if fuel[0] != fuel[1]:
        return S.WIN_A if fuel[0] > fuel[1] else S.WIN_B
    return S.WIN_TIE


def words_to_actions(ref, game, unit_words, tile_bytes):
    """Packed action tensors of one match -> reference Action list (canonical order)."""
    A = ref.actions
    out = []
    pos = 0
    for city in game.cities.values():
        for cell in city.city_cells:
            code = int(tile_bytes[pos])
            pos += 1
            team = cell.city_tile.team
            if code == S.TA_BUILD_WORKER:
                out.append(A.SpawnWorkerAction(team, None, cell.pos.x, cell.pos.y))
            elif code == S.TA_BUILD_CART:
                out.append(A.SpawnCartAction(team, None, cell.pos.x, cell.pos.y))
            elif code == S.TA_RESEARCH:
                out.append(A.ResearchAction(team, cell.pos.x, cell.pos.y, None))
    slot = 0
    for team in (0, 1):
        for unit in game.state["teamStates"][team]["units"].values():
            a = S.decode_unit_action(unit_words[slot])
            slot += 1
            k = a["kind"]
            if k == S.UA_MOVE:
                d = a["dir"]
                out.append(A.MoveAction(team, unit.id, S.DIR_CHARS[d] if d < 5 else "x"))
            elif k == S.UA_BCITY:
                out.append(A.SpawnCityAction(team, unit.id))
            elif k == S.UA_PILLAGE:
                out.append(A.PillageAction(team, unit.id))
            elif k == S.UA_TRANSFER:
                out.append(A.TransferAction(team, unit.id, "u_%d" % a["dest"], S.RES_NAMES[a["res"]], a["amount"]))
    return out


def commands_to_words(game, commands_by_team, ms):
    """Kaggle command strings of one turn -> packed action tensors (for replay fixtures).

    Returns (unit_words, tile_bytes, exact) where `exact` is False when the turn holds
    something the packed form cannot express (two commands for one entity, a command for
    an entity that does not exist, spawn commands out of canonical order that hit the
    unit cap) -- the caller then keeps the reference in charge for that turn.
    Command grammar: luxai2021/game/game.py:299-388.
    """
    size = ms.size
    nu, nt = int(ms.hdr["n_units"]), int(ms.hdr["n_tiles"])
    unit_words = np.zeros(len(ms.units), dtype=np.uint32)
    tile_bytes = np.zeros(len(ms.tiles), dtype=np.uint8)
    exact = True
    slot_of_unit = {}
    for i in range(nu):
        slot_of_unit[(int(ms.units["team_type"][i]) & 1, int(ms.units["id"][i]))] = i
    pos_of_tile = {int(ms.tiles[i]): i for i in range(nt)}
    dirs = {c: i for i, c in enumerate(S.DIR_CHARS)}
    spawn_seen = []
    for team, cmds in enumerate(commands_by_team):
        for cmd in cmds or []:
            p = cmd.split(" ")
            op = p[0]
            if op in ("m", "bcity", "p", "t"):
                key = (team, int(p[1].split("_")[1])) if p[1].startswith("u_") else None
                if key not in slot_of_unit:
                    exact = False
                    continue
                s = slot_of_unit[key]
                if unit_words[s] != 0:
                    exact = False
                    continue
                if op == "m":
                    unit_words[s] = S.encode_move(dirs.get(p[2], 7))
                elif op == "bcity":
                    unit_words[s] = S.UA_BCITY
                elif op == "p":
                    unit_words[s] = S.UA_PILLAGE
                else:
                    unit_words[s] = S.encode_transfer(int(p[2].split("_")[1]), S.RES_CODES[p[3]], int(p[4]))
            elif op in ("bw", "bc", "r"):
                x, y = int(p[1]), int(p[2])
                idx = y * size + x
                if not (0 <= x < size and 0 <= y < size) or idx not in pos_of_tile:
                    exact = False
                    continue
                t = pos_of_tile[idx]
                owner = ((int(ms.cells["flags"][idx]) >> 2) & 3) - 1
                if tile_bytes[t] != 0 or (op != "r" and owner != team):
                    exact = False
                    continue
                tile_bytes[t] = {"bw": S.TA_BUILD_WORKER, "bc": S.TA_BUILD_CART, "r": S.TA_RESEARCH}[op]
                if op != "r":
                    spawn_seen.append((team, t))
            else:
                exact = False
    # spawn commands are validated in list order against the unit cap (game.py:696-718);
    # the packed form validates them in canonical tile order
    for team in (0, 1):
        seq = [t for tm, t in spawn_seen if tm == team]
        if seq != sorted(seq):
            exact = False
    return unit_words, tile_bytes, exact
