"""Kaggle wire formats as state import / export.  Host side only.

  state_from_updates   the per-turn update lines a Kaggle engine sends to agents ("rp", "r", "u", "c", "ct",
                       "ccd", terminated by "D_DONE") -> packed MatchState.  Mirrors
                       Game.process_updates(updates, assign=True), luxai2021/game/game.py:159-262, applied to
                       an empty map -- including what that import leaves out or adds in the reference:
                       adjacent_city_tiles stay 0 (Cell.set_city_tile / City.add_city_tile :51-59, :52-57 do
                       not count neighbours), every "u" line bumps workersBuilt / cartsBuilt
                       (spawn_worker / spawn_cart :744-796), every "ct" line bumps cityTilesBuilt (:247),
                       the id counters do not move (explicit ids).  `recount_adjacency=True` gives the tiles
                       their true neighbour counts instead (the game's semantics rather than the importer's).
  state_from_replay    the same for step `step` of a Kaggle replay JSON (observation of player 0: updates
                       after the two id / size lines, id counters, turn).
  updates_from_state   packed MatchState -> the update lines (what the engine would send for this state),
                       in the order the reference's importer expects: research, resources (map.resources
                       order), units (team A then B), cities, city tiles (cities x city_cells), roads.
  commands_to_actions  Kaggle command texts of one team -> Action objects (action_from_command_low,
                       game.py:299-388): same grammar checks, same exceptions for malformed commands.

Batch import: [state_from_replay(rep, s) for ...] -> BatchedGame.load_states(...) fills a device batch with
late-game states (dense benchmarks, offline datasets) without the reference engine.
"""
import numpy as np

from . import actions as A
from . import state as S

_TYPE_CODES = {"wood": S.RES_WOOD, "coal": S.RES_COAL, "uranium": S.RES_URANIUM}


def _quarter(value):
    q = float(value) * 4.0
    if q != int(q):
        raise ValueError("cooldown / road level %r is not a multiple of 0.25" % (value,))
    return int(q)


def _num(uid):
    return int(str(uid).split("_")[1])


def state_from_updates(size, updates, turn=0, unit_id_count=0, city_id_count=0, caps=None, recount_adjacency=False):
    """Update lines -> MatchState (see the module docstring for the exact importer semantics)."""
    ms = S.MatchState(size, caps)
    h = ms.hdr
    h["turn"], h["unit_id_count"], h["city_id_count"], h["size"] = turn, unit_id_count, city_id_count, size
    return apply_updates(ms, updates, recount_adjacency=recount_adjacency)


def apply_updates(base, updates, recount_adjacency=False):
    """Game.process_updates(updates, assign=True) (game.py:159-262) applied to the packed state `base`
    (an empty map for Game.reset(updates=...) as AgentFromStdInOut uses it, agent.py:224-259); returns a new
    MatchState.  Units, cities, tiles and resource cells of `base` stay, in their order, and the lines add to
    them as the reference's dict / list appends do.  A line that re-declares a resource cell or a city tile
    that already exists would leave the reference with duplicate list entries the packed form cannot hold:
    ValueError."""
    size = base.size
    ms = base.copy()
    h = ms.hdr
    nu0, nc0, nt0, nr0 = int(h["n_units"]), int(h["n_cities"]), int(h["n_tiles"]), int(h["n_res"])
    units = ([], [])     # per team: [id, x, y, team_type, cd_q, wood, coal, uranium]
    for i in range(nu0):
        u = base.units[i]
        units[int(u["team_type"]) & 1].append([int(u["id"]), int(u["x"]), int(u["y"]), int(u["team_type"]), int(u["cd_q"]),
                                               int(u["wood"]), int(u["coal"]), int(u["uranium"])])
    cities = {}          # id string -> [team, fuel, [cell indices]]  (dict order == Game.cities order)
    slots = []
    for s in range(nc0):
        c = base.cities[s]
        rec = [int(c["team"]), int(c["fuel"]), []]
        cities["c_%d" % int(c["id"])] = rec
        slots.append(rec)
    for p in range(nt0):
        idx = int(base.tiles[p])
        slots[int(base.cells["city_slot"][idx])][2].append(idx)
    res_order = [int(x) for x in S.res_cell(base.res[:nr0])]
    res_seen, tile_seen = set(res_order), set(int(x) for x in base.tiles[:nt0])
    for line in updates or []:
        if line == "D_DONE":
            break
        p = line.split(" ")
        op = p[0]
        if op == "rp":
            h["research"][int(p[1])] = int(p[2])
        elif op == "r":
            x, y, amt = int(p[2]), int(p[3]), int(float(p[4]))
            idx = y * size + x
            if idx in res_seen:
                raise ValueError("resource cell (%d, %d) declared twice" % (x, y))
            res_seen.add(idx)
            ms.cells["amount"][idx] = amt
            ms.cells["flags"][idx] = (int(ms.cells["flags"][idx]) & ~3) | (_TYPE_CODES[p[1]] if amt > 0 else 0)
            res_order.append(idx)
        elif op == "u":
            kind, team = int(p[1]), int(p[2])
            rec = [_num(p[3]), int(p[4]), int(p[5]), team | (kind << 1), _quarter(p[6]), int(p[7]), int(p[8]), int(p[9])]
            for k, old in enumerate(units[team]):
                if old[0] == rec[0]:
                    units[team][k] = rec             # same id: the dict entry is overwritten in place
                    break
            else:
                units[team].append(rec)
            h["carts_built" if kind == 1 else "workers_built"][team] += 1
        elif op == "c":
            if float(p[3]) != int(float(p[3])):
                raise ValueError("fractional city fuel %r" % p[3])
            if p[2] in cities and cities[p[2]][2]:
                raise ValueError("city %s declared twice" % p[2])
            cities[p[2]] = [int(p[1]), int(float(p[3])), []]
        elif op == "ct":
            team, x, y = int(p[1]), int(p[3]), int(p[4])
            idx = y * size + x
            city = cities[p[2]]                      # KeyError like the reference (:241)
            cd = float(p[5])
            if cd != int(cd):
                raise ValueError("fractional city tile cooldown %r" % p[5])
            if idx in tile_seen:
                raise ValueError("city tile (%d, %d) declared twice" % (x, y))
            tile_seen.add(idx)
            ms.cells["flags"][idx] = (int(ms.cells["flags"][idx]) & 3) | ((team + 1) << 2)
            ms.cells["tile_cd"][idx] = int(cd)
            city[2].append(idx)
            h["city_tiles_built"][team] += 1
        elif op == "ccd":
            ms.cells["road_q"][int(p[2]) * size + int(p[1])] = _quarter(p[3])
    # resources list: GameMap.resources order, cells with amount > 0 (oracle/flatten.py keeps the same rule)
    ms.set_res_cells([idx for idx in res_order if int(ms.cells["amount"][idx]) > 0])
    # cities and tiles in dict order x city_cells order
    if len(cities) > len(ms.cities):
        raise ValueError("more cities than the capacity %d" % len(ms.cities))
    ntile = 0
    ms.cities[:] = 0
    ms.tiles[:] = 0
    h["n_team_tiles"][:] = 0
    for slot, (cid, (team, fuel, cells)) in enumerate(cities.items()):
        c = ms.cities[slot]
        c["id"], c["fuel"], c["team"], c["ntiles"] = _num(cid), fuel, team, len(cells)
        adjsum = 0
        for key, idx in enumerate(cells):
            if ntile >= len(ms.tiles):
                raise ValueError("more city tiles than the capacity %d" % len(ms.tiles))
            ms.tiles[ntile] = idx
            ntile += 1
            ms.cells["city_slot"][idx] = slot
            ms.cells["key"][idx] = key
            adjsum += (int(ms.cells["flags"][idx]) >> 4) & 7
        c["adjsum"] = adjsum
        h["n_team_tiles"][team] += len(cells)
    h["n_cities"], h["n_tiles"] = len(cities), ntile
    if recount_adjacency:
        for slot in range(len(cities)):
            adjsum = 0
            for p in range(ntile):
                idx = int(ms.tiles[p])
                if int(ms.cells["city_slot"][idx]) != slot:
                    continue
                team = ((int(ms.cells["flags"][idx]) >> 2) & 3) - 1
                x, y = idx % size, idx // size
                adj = 0
                for dx, dy in ((0, -1), (1, 0), (0, 1), (-1, 0)):
                    nx, ny = x + dx, y + dy
                    if 0 <= nx < size and 0 <= ny < size:
                        f = int(ms.cells["flags"][ny * size + nx])
                        adj += ((f >> 2) & 3) - 1 == team
                ms.cells["flags"][idx] = (int(ms.cells["flags"][idx]) & 0x0F) | (adj << 4)
                adjsum += adj
            ms.cities["adjsum"][slot] = adjsum
    # units: team A in line order, then team B
    k = 0
    ms.units[:] = 0
    for team in (0, 1):
        h["n_team_units"][team] = len(units[team])
        for rec in units[team]:
            if k >= len(ms.units):
                raise ValueError("more units than the capacity %d" % len(ms.units))
            u = ms.units[k]
            u["id"], u["x"], u["y"], u["team_type"], u["cd_q"], u["wood"], u["coal"], u["uranium"] = rec
            k += 1
    h["n_units"] = k
    return ms


def state_from_replay(replay, step=0, caps=None, recount_adjacency=False):
    """Step `step` of a Kaggle replay JSON (dict) -> MatchState, as the reference's test harness imports it
    (luxai2021/tests/test_replay.py: Game.reset(updates=obs["updates"][2:]) on an empty map)."""
    obs = replay["steps"][step][0]["observation"]
    size = int(obs["width"])
    if int(obs["height"]) != size:
        raise ValueError("square maps only")
    return state_from_updates(size, obs["updates"][2:], turn=int(obs["step"]),
                              unit_id_count=int(obs["globalUnitIDCount"]), city_id_count=int(obs["globalCityIDCount"]),
                              caps=caps, recount_adjacency=recount_adjacency)


def _fmt(q):
    """Quarter units -> the shortest decimal text (2 -> "0.5", 8 -> "2")."""
    v = q / 4.0
    return str(int(v)) if v == int(v) else repr(v)


def updates_from_state(ms):
    """MatchState -> update lines (inverse of state_from_updates up to the fields the wire format does not
    carry: statistics, id counters, turn, adjacent_city_tiles)."""
    size, h = ms.size, ms.hdr
    out = ["rp 0 %d" % int(h["research"][0]), "rp 1 %d" % int(h["research"][1])]
    names = S.RES_NAMES
    for r in range(int(h["n_res"])):
        idx = int(S.res_cell(ms.res[r]))
        amt = int(ms.cells["amount"][idx])
        if amt > 0:
            out.append("r %s %d %d %d" % (names[int(ms.cells["flags"][idx]) & 3], idx % size, idx // size, amt))
    for i in range(int(h["n_units"])):
        u = ms.units[i]
        tt = int(u["team_type"])
        out.append("u %d %d u_%d %d %d %s %d %d %d" % (tt >> 1, tt & 1, int(u["id"]), int(u["x"]), int(u["y"]),
                                                       _fmt(int(u["cd_q"])), int(u["wood"]), int(u["coal"]), int(u["uranium"])))
    for s in range(int(h["n_cities"])):
        c = ms.cities[s]
        upkeep = 23 * int(c["ntiles"]) - 5 * int(c["adjsum"])
        out.append("c %d c_%d %d %d" % (int(c["team"]), int(c["id"]), int(c["fuel"]), upkeep))
    for p in range(int(h["n_tiles"])):
        idx = int(ms.tiles[p])
        cell = ms.cells[idx]
        team = ((int(cell["flags"]) >> 2) & 3) - 1
        out.append("ct %d c_%d %d %d %d" % (team, int(ms.cities["id"][int(cell["city_slot"])]), idx % size, idx // size,
                                            int(cell["tile_cd"])))
    for idx in np.nonzero(ms.cells["road_q"])[0]:
        out.append("ccd %d %d %s" % (int(idx) % size, int(idx) // size, _fmt(int(ms.cells["road_q"][idx]))))
    # city tiles report the effective road level 6 (cell.py:77-85) like the engine does
    for p in range(int(h["n_tiles"])):
        idx = int(ms.tiles[p])
        if int(ms.cells["road_q"][idx]) == 0:
            out.append("ccd %d %d 6" % (idx % size, idx // size))
    out.append("D_DONE")
    return out


def commands_to_actions(commands, team, game=None):
    """Kaggle command texts of one team -> Action list (action_from_command_low, game.py:299-388).
    Malformed commands raise like the reference; with `game` (a luxpythonenvgym_b200.game.Game) unit ids of
    pillage / build-city commands are checked against the team's units as the reference does (:317-330)."""
    out = []
    for comm in commands:
        parts = comm.split(" ")
        if len(parts) <= 1:
            raise Exception("Agent %d sent invalid command; cmd: %s" % (team, comm))
        op, args = parts[0], parts[1:]
        want = {"p": 1, "bcity": 1, "bc": 2, "bw": 2, "m": 2, "r": 2, "t": 4}.get(op)
        if want is None:
            raise Exception("unknown action %s" % op)
        if len(args) != want:
            raise Exception("Agent %d sent malformed command: %s" % (team, comm))
        if game is not None and op in ("p", "bcity"):
            game.get_unit(team, args[0])             # KeyError for a unit the team does not have
        out.append(A.action_from_string(comm, team))
    return out
