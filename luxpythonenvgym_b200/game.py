"""`Game`: the reference's single-match API (luxai2021/game/game.py) on top of the batched CUDA engine.

`Game(configs)`, `.reset(updates=None, increment_turn=False)`, `.run_turn_with_actions(actions) -> bool`
(True = match over), `.process_updates(updates, assign)`, and the object model agents look at (`.state`,
`.stats`, `.cities`, `.map`, `.get_unit`, ...), refreshed from the device state after every turn.  Unit, city
tile and cell objects keep their identity across turns like the reference's (MatchController keys action
sequences by them).  It is the drop-in adapter for existing per-match code -- `MatchController`,
`LuxEnvironment`, agents; throughput comes from `BatchedGame` / `BatchedLuxEnvironment`, not from here.
The turn always runs on the device: there is no CPU engine behind this class.

List semantics of `run_turn_with_actions(actions)` (game.py:390-465), reproduced on the host before the packed
action words are written:
  * every action goes through validate_command (:646-723) IN LIST ORDER with the per-team spawn counters, so
    when the unit cap binds the same spawns survive as in the reference; rejected actions are dropped like
    the reference drops them (:422-425);
  * an entity that is left holding several validated actions does nothing (unit.py:172, city.py:101).  The
    one thing not reproduced for such an entity: its extra MOVES do not take part in the collision pass
    (the packed engine holds one action per entity); `dropped_actions` counts them;
  * a transfer whose destination does not exist or does not parse is the reference's swallowed KeyError:
    no transfer, no cooldown.
"""
import numpy as np

from . import actions as A
from . import mapgen
from . import state as S
from . import wire
from .batched import BatchedGame
from .constants import Constants, LuxMatchConfigs_Default, GAME_CONSTANTS

DEFAULT_CONFIGS = LuxMatchConfigs_Default


class MatchWarn(Exception):
    pass


class Position:
    """luxai2021/game/position.py"""

    def __init__(self, x, y):
        self.x, self.y = x, y

    def __sub__(self, pos):
        return abs(pos.x - self.x) + abs(pos.y - self.y)

    distance_to = __sub__

    def is_adjacent(self, pos):
        return (self - pos) <= 1

    def __eq__(self, pos):
        return pos is not None and self.x == pos.x and self.y == pos.y

    equals = __eq__

    def __hash__(self):
        return 10000 * self.x + self.y

    def translate(self, direction, units):
        dx, dy = {"n": (0, -1), "e": (1, 0), "s": (0, 1), "w": (-1, 0), "c": (0, 0)}[direction]
        return Position(self.x + dx * units, self.y + dy * units)

    def direction_to(self, target):
        best, direction = self.distance_to(target), "c"
        for d in ("n", "e", "s", "w"):
            dist = target.distance_to(self.translate(d, 1))
            if dist < best:
                best, direction = dist, d
        return direction

    def __str__(self):
        return "({}, {})".format(self.x, self.y)


class Resource:
    def __init__(self, resource_type, amount):
        self.type, self.amount = resource_type, amount


class _Actionable:
    """actionable.py:27-43: `can_act` with MatchController's per-turn override."""
    cooldown = 0.0
    can_act_override = None

    def can_act(self):
        return self.cooldown < 1 if self.can_act_override is None else self.can_act_override

    def set_can_act_override(self, can_act_override):
        self.can_act_override = can_act_override


class Unit(_Actionable):
    def __init__(self, uid, team, unit_type):
        self.id, self.team, self.type = uid, team, unit_type
        self.pos = Position(0, 0)
        self.cargo = {"wood": 0, "uranium": 0, "coal": 0}

    def _load(self, x, y, cd_q, wood, coal, uranium):
        self.pos = Position(x, y)
        self.cooldown = cd_q / 4.0
        self.cargo["wood"], self.cargo["coal"], self.cargo["uranium"] = wood, coal, uranium

    def is_worker(self):
        return self.type == 0

    def is_cart(self):
        return self.type == 1

    def can_move(self):
        return self.can_act()

    def get_cargo_space_left(self):
        return (2000 if self.type else 100) - (self.cargo["wood"] + self.cargo["coal"] + self.cargo["uranium"])

    def get_cargo_fuel_value(self):
        return self.cargo["wood"] + 10 * self.cargo["coal"] + 40 * self.cargo["uranium"]

    def get_light_upkeep(self):
        return 10 if self.type else 4

    def can_build(self, game_map):
        """unit.py:100-109"""
        cell = game_map.get_cell_by_pos(self.pos)
        return (not cell.has_resource()) and self.can_act() and \
            self.cargo["wood"] + self.cargo["coal"] + self.cargo["uranium"] >= 100


class Worker(Unit):
    pass


class Cart(Unit):
    pass


class CityTile(_Actionable):
    def __init__(self, team, pos, city_id):
        self.team, self.pos, self.city_id = team, pos, city_id
        self.adjacent_city_tiles = 0

    def can_build_unit(self):
        return self.can_act()

    can_research = can_build_unit

    def get_tile_id(self):
        return "{}_{}_{}".format(self.city_id, self.pos.x, self.pos.y)

    def get_cargo_space_left(self):
        return 9999999


class City:
    def __init__(self, cid, team):
        self.id, self.team, self.fuel = cid, team, 0
        self.city_cells = []
        self._upkeep = 0

    def get_light_upkeep(self):
        return self._upkeep

    def get_adjacency_bonuses(self):
        return 5 * sum(cell.city_tile.adjacent_city_tiles for cell in self.city_cells)


class Cell:
    def __init__(self, x, y):
        self.pos = Position(x, y)
        self.resource, self.city_tile, self.units, self.road = None, None, {}, 0.0

    def has_resource(self):
        return self.resource is not None and self.resource.amount > 0

    def is_city_tile(self):
        return self.city_tile is not None

    def has_units(self):
        return len(self.units) != 0

    def get_road(self):
        return 6 if self.city_tile is not None else self.road


class GameMap:
    def __init__(self, size):
        self.width = self.height = size
        self.map = [[Cell(x, y) for x in range(size)] for y in range(size)]
        self.resources = []
        self.resources_by_type = {"wood": [], "coal": [], "uranium": []}

    def in_map(self, pos):
        return 0 <= pos.x < self.width and 0 <= pos.y < self.height

    def get_cell(self, x, y):
        return self.map[y][x] if 0 <= x < self.width and 0 <= y < self.height else None

    def get_cell_by_pos(self, pos):
        return self.get_cell(pos.x, pos.y)

    def get_row(self, y):
        return self.map[y]

    def get_adjacent_cells(self, cell):
        x, y = cell.pos.x, cell.pos.y
        out = []
        if y > 0:
            out.append(self.map[y - 1][x])
        if x < self.width - 1:
            out.append(self.map[y][x + 1])
        if y < self.height - 1:
            out.append(self.map[y + 1][x])
        if x > 0:
            out.append(self.map[y][x - 1])
        return out

    def get_adjacent_cells_with_corners(self, cell):
        out = self.get_adjacent_cells(cell)
        for dx, dy in ((-1, -1), (1, -1), (-1, 1), (1, 1)):
            c = self.get_cell(cell.pos.x + dx, cell.pos.y + dy)
            if c:
                out.append(c)
        return out

    def get_map_string(self):
        rows = []
        for row in self.map:
            s = ""
            for cell in row:
                if cell.has_units():
                    units = list(cell.units.values())
                    head = ("c" if units[0].type else "W") if len(units) == 1 else str(len(units))
                    s += head + "ab"[units[0].team]
                elif cell.has_resource():
                    s += cell.resource.type[0] + ","
                elif cell.is_city_tile():
                    s += "C" + "ab"[cell.city_tile.team]
                else:
                    s += ".."
            rows.append(s)
        return "\n".join(rows) + "\n"

    def to_state_object(self):
        grid = []
        for row in self.map:
            out = []
            for cell in row:
                d = {}
                if cell.get_road() != 0:
                    d["road"] = cell.get_road()
                if cell.resource:
                    d["type"], d["amount"] = cell.resource.type, cell.resource.amount
                out.append(d)
            grid.append(out)
        return grid


class Game:
    def __init__(self, configs=None, agents=None, device="cuda:0"):
        self.configs = dict(LuxMatchConfigs_Default)
        self.configs.update(configs or {})
        if self.configs.get("parameters", GAME_CONSTANTS["PARAMETERS"]) != GAME_CONSTANTS["PARAMETERS"]:
            raise ValueError("the CUDA engine is compiled for the default game parameters "
                             "(luxai2021/game/game_constants.json); configs['parameters'] differs")
        self.agents = []
        self.device = device
        self._batch = None
        self._ms_cache = None
        self._map = None
        self._stale = False
        self._state = {"turn": 0}
        self.dropped_actions = 0
        self.replay = None
        self.replay_stateful = self.replay_folder = self.replay_filename_prefix = None
        self.log_file = None
        self.reset()

    # ------------------------------------------------------------------ the lazily refreshed view
    # Code that drives the device batch directly (LuxEnvironment's all-device fast path) calls
    # `_mark_stale()` after a turn; the object model is rebuilt the next time anybody looks at it.
    def _mark_stale(self):
        self._stale = True

    def _sync(self):
        if self._stale:
            self._refresh()

    @property
    def _ms(self):
        self._sync()
        return self._ms_cache

    @property
    def state(self):
        self._sync()
        return self._state

    @property
    def stats(self):
        self._sync()
        return self._stats

    @property
    def cities(self):
        self._sync()
        return self._cities

    @property
    def map(self):
        self._sync()
        return self._map

    def _poke_header(self, field, value):
        ms = self._ms.copy()
        ms.hdr[field] = value
        self._batch.load_states([ms])
        self._refresh()

    @property
    def global_unit_id_count(self):
        return int(self._ms.hdr["unit_id_count"])

    @global_unit_id_count.setter
    def global_unit_id_count(self, value):
        self._poke_header("unit_id_count", int(value))

    @property
    def global_city_id_count(self):
        return int(self._ms.hdr["city_id_count"])

    @global_city_id_count.setter
    def global_city_id_count(self, value):
        self._poke_header("city_id_count", int(value))

    # ------------------------------------------------------------------ replay logging (game.py:38-74)
    def start_replay_logging(self, stateful=False, replay_folder="./replays/", replay_filename_prefix="replay"):
        import os
        import random
        from . import replay as R
        assert self.configs.get("seed") is not None, "Replays only work when a seed is specified."
        os.makedirs(replay_folder, exist_ok=True)
        filename = "%s_%d.json" % (replay_filename_prefix, random.randint(0, 10000))
        self.replay = R.Replay(self.configs["seed"], self._batch.size, stateful=stateful)
        self.replay.path = os.path.join(replay_folder, filename)
        self.replay_stateful, self.replay_folder, self.replay_filename_prefix = stateful, replay_folder, replay_filename_prefix

    def stop_replay_logging(self):
        self.replay = None
        self.replay_stateful = self.replay_folder = self.replay_filename_prefix = None

    # ------------------------------------------------------------------ lifecycle (game.py:76-157)
    def reset(self, updates=None, increment_turn=False, state=None):
        """A new match: the seeded map of configs["seed"] (configs["width"] overrides the size), an empty
        map with configs["mapType"] == "empty", or a packed MatchState passed as `state` (an extension).
        `updates` (Kaggle update lines) are imported on top, exactly Game.process_updates(assign=True);
        `increment_turn` keeps the turn counter and adds one (game.py:120-123)."""
        if state is not None and not isinstance(state, S.MatchState):
            raise TypeError("reset(state=...) takes a packed MatchState; update lines go to reset(updates=...)")
        turn = int(self.state["turn"]) + 1 if increment_turn else 0
        if isinstance(updates, S.MatchState):          # reset(<MatchState>) of the first version of this adapter
            state, updates = updates, None
        if state is None:
            if self.configs.get("mapType") == Constants.MAP_TYPES.EMPTY:
                size = self.configs.get("width")
                if size is None or self.configs.get("height", size) != size:
                    raise ValueError("an empty map needs configs['width'] == configs['height']")
                state = S.MatchState(int(size))
            else:
                seed = self.configs.get("seed")
                if seed is None:
                    seed = int(np.random.randint(0, 2 ** 31 - 1))
                if self.configs.get("height", self.configs.get("width")) != self.configs.get("width"):
                    raise ValueError("square maps only")
                state = mapgen.generate(int(seed), size=self.configs.get("width"))
        if self._batch is None or self._batch.size != state.size:
            self._batch = BatchedGame(1, state.size, device=self.device)
            self._map = None
        ms = S.MatchState(state.size, self._batch.caps)
        ms.hdr, ms.cells = state.hdr.copy(), state.cells.copy()
        for name in ("units", "cities", "tiles", "res"):
            src, dst = getattr(state, name), getattr(ms, name)
            k = min(len(src), len(dst))
            dst[:k] = src[:k]
        ms.hdr["turn"] = turn
        if updates is not None:
            ms = wire.apply_updates(ms, updates)
        self._units_by_id = ({}, {})
        self._tiles_by_cell = {}
        self._batch.load_states([ms])
        if self.replay:
            self.replay.clear()
        self._refresh()

    def process_updates(self, updates, assign=True):
        """game.py:159-262.  assign=True imports the update lines into the running match (state on the
        device replaced); assign=False asserts that the match equals them, field by field as the reference
        does (research points and flags, resource type and amount per cell, every unit id on its cell, city
        ids, city-tile membership, road levels)."""
        if updates is None:
            return
        if assign:
            ms = wire.apply_updates(self._ms, updates)
            self._batch.load_states([ms])
            self._refresh()
            return
        ts = self.state["teamStates"]
        for line in updates:
            if line == "D_DONE":
                break
            p = line.split(" ")
            op = p[0]
            if op == "rp":
                team, rp = int(p[1]), int(p[2])
                assert ts[team]["researchPoints"] == rp
                if rp >= 50:
                    assert ts[team]["researched"]["coal"] is True
                if rp >= 200:
                    assert ts[team]["researched"]["uranium"] is True
            elif op == "r":
                cell = self.map.get_cell(int(p[2]), int(p[3]))
                assert cell.resource.amount == int(float(p[4]))
                assert cell.resource.type == p[1]
            elif op == "u":
                cell = self.map.get_cell(int(p[4]), int(p[5]))
                assert len(cell.units) > 0
                assert p[3] in [u.id for u in cell.units.values()], "unit id %s missplaced" % p[3]
            elif op == "c":
                assert p[2] in self.cities
            elif op == "ct":
                city = self.cities[p[2]]
                cell = self.map.get_cell(int(p[3]), int(p[4]))
                assert cell.city_tile.city_id == p[2]
                assert cell in city.city_cells
            elif op == "ccd":
                assert self.map.get_cell(int(p[1]), int(p[2])).get_road() == float(p[3])

    def _refresh(self):
        """Object model <- device state.  One D2H copy of the match; objects are updated in place."""
        self._stale = False
        ms = self._batch.export_states([0])[0]
        self._ms_cache = ms
        size = ms.size
        h = ms.hdr
        if self._map is None or self._map.width != size:
            self._map = GameMap(size)
            self._road_q = np.zeros(size * size, dtype=np.uint8)
            self._occupied, self._res_cells, self._tile_cells = [], [], []
        gmap = self._map
        grid = gmap.map
        n_units, n_cities, n_tiles, n_res = int(h["n_units"]), int(h["n_cities"]), int(h["n_tiles"]), int(h["n_res"])
        cells = ms.cells
        flags, amount = cells["flags"], cells["amount"]
        # roads: only the cells whose level changed
        rq = cells["road_q"]
        for idx in np.nonzero(rq != self._road_q)[0].tolist():
            grid[idx // size][idx % size].road = int(rq[idx]) / 4.0
        self._road_q = rq.copy()
        # resources, in GameMap.resources order (depleted cells leave the list, game.py:498-510)
        for cell in self._res_cells:
            cell.resource = None
        gmap.resources = []
        gmap.resources_by_type = {"wood": [], "coal": [], "uranium": []}
        self._res_cells = []
        res_idx = S.res_cell(ms.res[:n_res]).tolist()
        for idx, amt, fl in zip(res_idx, amount[res_idx].tolist(), flags[res_idx].tolist()):
            if amt > 0:
                cell = grid[idx // size][idx % size]
                cell.resource = Resource(S.RES_NAMES[fl & 3], amt)
                gmap.resources.append(cell)
                gmap.resources_by_type[cell.resource.type].append(cell)
                self._res_cells.append(cell)
        # cities and their tiles (cities x city_cells order)
        self._cities = {}
        slots = []
        crec = ms.cities[:n_cities]
        for cid, team, fuel, nt, adj in zip(crec["id"].tolist(), crec["team"].tolist(), crec["fuel"].tolist(),
                                            crec["ntiles"].tolist(), crec["adjsum"].tolist()):
            city = City("c_%d" % cid, team)
            city.fuel, city._upkeep = fuel, 23 * nt - 5 * adj
            self._cities[city.id] = city
            slots.append(city)
        for cell in self._tile_cells:
            cell.city_tile = None
        self._tile_cells = []
        fresh = {}
        tidx = ms.tiles[:n_tiles].tolist()
        for idx, fl, cd, slot in zip(tidx, flags[tidx].tolist(), cells["tile_cd"][tidx].tolist(),
                                     cells["city_slot"][tidx].tolist()):
            cell = grid[idx // size][idx % size]
            city = slots[slot]
            team = ((fl >> 2) & 3) - 1
            tile = self._tiles_by_cell.get(idx)
            if tile is None or tile.team != team:
                tile = CityTile(team, cell.pos, city.id)
            tile.city_id, tile.cooldown, tile.adjacent_city_tiles = city.id, float(cd), (fl >> 4) & 7
            tile.can_act_override = None
            fresh[idx] = tile
            cell.city_tile = tile
            city.city_cells.append(cell)
            self._tile_cells.append(cell)
        self._tiles_by_cell = fresh
        # units: team A in dict order, then team B
        for cell in self._occupied:
            cell.units = {}
        self._occupied = []
        team_states = {}
        for t in (0, 1):
            rp = int(h["research"][t])
            team_states[t] = {"researchPoints": rp, "units": {},
                              "researched": {"wood": True, "coal": rp >= 50, "uranium": rp >= 200}}
        urec = ms.units[:n_units]
        by_id = ({}, {})
        self._slot_of = {}
        for i, (uid, x, y, tt, cd, w, c, u) in enumerate(zip(
                urec["id"].tolist(), urec["x"].tolist(), urec["y"].tolist(), urec["team_type"].tolist(),
                urec["cd_q"].tolist(), urec["wood"].tolist(), urec["coal"].tolist(), urec["uranium"].tolist())):
            team, kind = tt & 1, (tt >> 1) & 1
            unit = self._units_by_id[team].get(uid)
            if unit is None or unit.type != kind:
                unit = (Cart if kind else Worker)("u_%d" % uid, team, kind)
            unit._load(x, y, cd, w, c, u)
            unit.can_act_override = None
            by_id[team][uid] = unit
            team_states[team]["units"][unit.id] = unit
            cell = grid[y][x]
            if not cell.units:
                self._occupied.append(cell)
            cell.units[unit.id] = unit
            self._slot_of[(team, unit.id)] = i
        self._units_by_id = by_id
        self._state = {"turn": int(h["turn"]), "teamStates": team_states}
        self._stats = {"teamStats": {t: {
            "fuelGenerated": int(h["fuel_generated"][t]),
            "resourcesCollected": {n: int(h["collected"][t][k]) for k, n in enumerate(("wood", "coal", "uranium"))},
            "cityTilesBuilt": int(h["city_tiles_built"][t]), "workersBuilt": int(h["workers_built"][t]),
            "cartsBuilt": int(h["carts_built"][t]), "roadsBuilt": int(h["roads_built_q"][t]) / 4.0,
            "roadsPillaged": 0} for t in (0, 1)}}
        self.cells_with_roads = None

    # ------------------------------------------------------------------ validation (game.py:646-723)
    def _gen_initial_accumulated_action_stats(self):
        return {t: {"workersBuilt": 0, "cartsBuilt": 0, "actionsPlaced": set()} for t in (0, 1)}

    def validate_command(self, cmd, accumulated_action_stats=None):
        """The action if the engine will accept it; raises MatchWarn / KeyError / AttributeError like the
        reference for what it drops."""
        if accumulated_action_stats is None:
            accumulated_action_stats = self._gen_initial_accumulated_action_stats()
        acc = accumulated_action_stats[cmd.team]
        if isinstance(cmd, A.SpawnCityAction):
            unit = self.get_unit(cmd.team, cmd.unit_id)
            cell = self.map.get_cell_by_pos(unit.pos)
            if cell.is_city_tile():
                raise MatchWarn("Agent tried to build CityTile on existing CityTile")
            if cell.has_resource():
                raise MatchWarn("Agent tried to build CityTile on non-empty resource tile")
            if not unit.cooldown < 1:
                raise MatchWarn("Agent tried to build CityTile with cooldown: {}".format(unit.cooldown))
            if sum(unit.cargo.values()) < 100:
                raise MatchWarn("Agent tried to build CityTile with insufficient materials")
            acc["actionsPlaced"].add(cmd.unit_id)
        elif isinstance(cmd, A.MoveAction):
            unit = self.get_unit(cmd.team, cmd.unit_id)
            if not unit.cooldown < 1:
                raise MatchWarn("Agent tried to move unit {} with cooldown: {}".format(cmd.unit_id, unit.cooldown))
            if cmd.direction not in ("c", "e", "n", "s", "w"):
                raise MatchWarn("Agent tried to move unit {} in invalid direction {}".format(cmd.unit_id, cmd.direction))
            if cmd.direction != "c":
                target = unit.pos.translate(cmd.direction, 1)
                if not self.map.in_map(target):
                    raise MatchWarn("Agent tried to move unit {} off map".format(cmd.unit_id))
                cell = self.map.get_cell_by_pos(target)
                if cell.is_city_tile() and cell.city_tile.team != cmd.team:
                    raise MatchWarn("Agent tried to move unit {} onto opponent CityTile".format(cmd.unit_id))
            acc["actionsPlaced"].add(cmd.unit_id)
        elif isinstance(cmd, A.SpawnAction):
            if not self.map.in_map(Position(cmd.x, cmd.y)):
                raise MatchWarn("Agent tried to build unit with invalid coordinates")
            tile = self.map.get_cell(cmd.x, cmd.y).city_tile
            if tile.team != cmd.team:                      # AttributeError off a city tile, like the reference
                raise MatchWarn("Agent tried to build unit on tile ({}, {}) that it does not own".format(cmd.x, cmd.y))
            if not tile.cooldown < 1:
                raise MatchWarn("Agent tried to build unit on tile ({}, {}) but CityTile still with cooldown".format(cmd.x, cmd.y))
            if self.worker_unit_cap_reached(cmd.team, acc["cartsBuilt"] + acc["workersBuilt"]):
                raise MatchWarn("Agent tried to build unit on tile ({}, {}) but unit cap reached".format(cmd.x, cmd.y))
            acc["actionsPlaced"].add(tile.get_tile_id())
            acc["cartsBuilt" if isinstance(cmd, A.SpawnCartAction) else "workersBuilt"] += 1
        return cmd

    # ------------------------------------------------------------------ the turn (game.py:390-535)
    def run_turn_with_actions(self, actions):
        ms = self._ms
        size = ms.size
        acc = self._gen_initial_accumulated_action_stats()
        held_units, held_tiles = {}, {}               # slot / tile position -> validated actions, list order
        tile_pos = None
        for act in actions or []:
            if act is None:
                continue
            try:
                # the validation itself looks at the cooldown, not at MatchController's override
                act = self.validate_command(act, acc)
            except Exception:
                self.dropped_actions += 1
                continue
            if isinstance(act, (A.SpawnAction, A.ResearchAction)):
                if tile_pos is None:
                    tile_pos = {int(ms.tiles[p]): p for p in range(int(ms.hdr["n_tiles"]))}
                inside = 0 <= act.x < size and 0 <= act.y < size
                p = tile_pos.get(act.y * size + act.x) if inside else None
                if p is None:
                    # research on a cell that is no city tile raises out of the reference's turn (game.py:447-450)
                    raise AttributeError("'NoneType' object has no attribute 'give_action'")
                held_tiles.setdefault(p, []).append(act)
            else:
                slot = self._slot_of.get((act.team, act.unit_id))
                if slot is None:
                    # pillage / transfer for a unit that does not exist raises out of the reference's turn (:442-453)
                    raise KeyError(act.unit_id)
                held_units.setdefault(slot, []).append(act)
        uw = np.zeros(len(ms.units), dtype=np.uint32)
        tb = np.zeros(len(ms.tiles), dtype=np.uint8)
        for p, held in held_tiles.items():
            if len(held) != 1:
                self.dropped_actions += len(held)
                continue
            act = held[0]
            tb[p] = (S.TA_RESEARCH if isinstance(act, A.ResearchAction)
                     else S.TA_BUILD_CART if isinstance(act, A.SpawnCartAction) else S.TA_BUILD_WORKER)
        for slot, held in held_units.items():
            # CENTER moves are never handed to the unit (game.py:462-465)
            real = [a for a in held if not (isinstance(a, A.MoveAction) and a.direction == "c")]
            if len(real) == 1 and len(held) == 1:
                uw[slot] = real[0].word()
            elif len(held) == 1:
                uw[slot] = held[0].word()             # a lone CENTER move still takes part in the collision pass
            elif len(real) == 1 and all(isinstance(a, A.MoveAction) for a in held if a is not real[0]):
                # one real action plus CENTER moves: the unit holds exactly one action after pruning
                uw[slot] = real[0].word()
                self.dropped_actions += len(held) - 1
            else:
                self.dropped_actions += len(held)
        if self.replay:
            self.replay.add_turn(ms, uw, tb)
        self._batch.put_actions(0, uw, tb)
        self._batch.step()
        self._refresh()
        over = bool(self._ms.hdr["done"])
        if self.replay:
            if self.replay.stateful:
                self.replay.add_state(self._ms)
            if over:
                self.replay.set_result(int(self._ms.hdr["winner"]))
                self.replay.write(self.replay.path)
        return over

    # ------------------------------------------------------------------ queries
    def get_teams_units(self, team):
        return self.state["teamStates"][team]["units"]

    def get_unit(self, team, unit_id):
        return self.state["teamStates"][team]["units"][unit_id]

    def is_night(self):
        return self.state["turn"] % 40 >= 30

    def match_over(self):
        if self.state["turn"] >= 359 or self._ms.hdr["done"]:
            return True
        cities = [0, 0]
        for c in self.cities.values():
            cities[c.team] += 1
        return any(len(self.get_teams_units(t)) + cities[t] == 0 for t in (0, 1))

    def get_winning_team(self):
        """Team index; ties that the reference breaks with a host coin flip (game.py:632) return team 1
        only after comparing city tiles, units and fuel -- check `hdr.winner == WIN_TIE` for the tie."""
        tiles = [sum(len(c.city_cells) for c in self.cities.values() if c.team == t) for t in (0, 1)]
        units = [len(self.get_teams_units(t)) for t in (0, 1)]
        fuel = [self.stats["teamStats"][t]["fuelGenerated"] for t in (0, 1)]
        for a in (tiles, units, fuel):
            if a[0] != a[1]:
                return 0 if a[0] > a[1] else 1
        return 1

    def worker_unit_cap_reached(self, team, offset=0):
        tiles = sum(len(c.city_cells) for c in self.cities.values() if c.team == team)
        return len(self.get_teams_units(team)) + offset >= tiles

    cart_unit_cap_reached = worker_unit_cap_reached

    def action_from_string(self, command, team):
        """game.py:292-297: a command naming a unit the team does not have gives None."""
        try:
            return wire.commands_to_actions([command], team, game=self)[0]
        except KeyError:
            return None

    def action_from_command_low(self, command, team):
        return wire.commands_to_actions([command], team, game=self)[0]

    def log(self, text):
        pass

    def packed_state(self):
        """The packed MatchState (host copy) of the current turn."""
        return self._ms.copy()

    def to_state_object(self):
        cities = {c.id: {"id": c.id, "fuel": c.fuel, "lightupkeep": c.get_light_upkeep(), "team": c.team,
                         "cityCells": [{"x": cell.pos.x, "y": cell.pos.y, "cooldown": cell.city_tile.cooldown}
                                       for cell in c.city_cells]} for c in self.cities.values()}
        teams = {}
        for t in (0, 1):
            ts = self.state["teamStates"][t]
            teams[t] = {"researchPoints": ts["researchPoints"], "researched": dict(ts["researched"]),
                        "units": {u.id: {"cargo": dict(u.cargo), "cooldown": u.cooldown, "x": u.pos.x, "y": u.pos.y,
                                         "type": u.type} for u in ts["units"].values()}}
        return {"turn": self.state["turn"], "globalCityIDCount": self.global_city_id_count,
                "globalunit_idCount": self.global_unit_id_count, "teamStates": teams,
                "map": self.map.to_state_object(), "cities": cities}
