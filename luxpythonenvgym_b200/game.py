"""`Game`: the reference's single-match API (luxai2021/game/game.py) on top of the batched engine.

`Game(configs)`, `.reset()`, `.run_turn_with_actions(actions) -> bool` (True = match over), and the
read-only object model agents look at (`.state`, `.stats`, `.cities`, `.map`, `.get_unit`, ...), all
materialised from the device state after every turn.  It is the drop-in adapter for existing
per-match code; throughput comes from `BatchedGame` / `BatchedLuxEnvironment`, not from here.

Domain notes: one action per entity per turn (extra ones for the same entity are
dropped and counted in `.dropped_actions`), actions for entities that do not exist are dropped like the
reference drops them (game.py:422-425), spawn actions are validated in cities x city_cells order.
"""
import numpy as np

from . import actions as A
from . import mapgen
from . import state as S
from .batched import BatchedGame

DEFAULT_CONFIGS = {"mapType": "random", "seed": None, "storeReplay": True, "debug": False}


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


class Unit:
    def __init__(self, rec):
        self.id = "u_%d" % int(rec["id"])
        self.team = int(rec["team_type"]) & 1
        self.type = (int(rec["team_type"]) >> 1) & 1
        self.pos = Position(int(rec["x"]), int(rec["y"]))
        self.cooldown = int(rec["cd_q"]) / 4.0
        self.cargo = {"wood": int(rec["wood"]), "uranium": int(rec["uranium"]), "coal": int(rec["coal"])}

    def is_worker(self):
        return self.type == 0

    def is_cart(self):
        return self.type == 1

    def can_act(self):
        return self.cooldown < 1

    can_move = can_act

    def get_cargo_space_left(self):
        return (2000 if self.type else 100) - sum(self.cargo.values())


class CityTile:
    def __init__(self, team, pos, city_id, cooldown, adjacent):
        self.team, self.pos, self.city_id = team, pos, city_id
        self.cooldown, self.adjacent_city_tiles = float(cooldown), adjacent

    def can_act(self):
        return self.cooldown < 1

    can_build_unit = can_research = can_act

    def get_tile_id(self):
        return "{}_{}_{}".format(self.city_id, self.pos.x, self.pos.y)


class City:
    def __init__(self, rec):
        self.id = "c_%d" % int(rec["id"])
        self.team, self.fuel = int(rec["team"]), int(rec["fuel"])
        self._upkeep = 23 * int(rec["ntiles"]) - 5 * int(rec["adjsum"])
        self.city_cells = []

    def get_light_upkeep(self):
        return self._upkeep


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
        self.resources, self.resources_by_type = [], {}

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


class Game:
    def __init__(self, configs=None, device="cuda:0"):
        self.configs = dict(DEFAULT_CONFIGS)
        self.configs.update(configs or {})
        self.device = device
        self._batch = None
        self.dropped_actions = 0
        self.reset()

    # ------------------------------------------------------------------ lifecycle
    def reset(self, state=None, **_ignored):
        """New match from configs["seed"] (+ optional "width"), or from a packed MatchState."""
        if state is None:
            seed = self.configs.get("seed")
            if seed is None:
                seed = int(np.random.randint(0, 2 ** 31 - 1))
            state = mapgen.generate(seed, size=self.configs.get("width"))
        if self._batch is None or self._batch.size != state.size:
            self._batch = BatchedGame(1, state.size, device=self.device)
        caps = self._batch.caps
        ms = S.MatchState(state.size, caps)
        ms.hdr, ms.cells = state.hdr.copy(), state.cells.copy()
        for name in ("units", "cities", "tiles", "res"):
            src, dst = getattr(state, name), getattr(ms, name)
            k = min(len(src), len(dst))
            dst[:k] = src[:k]
        self._batch.load_states([ms])
        self._refresh()

    def _refresh(self):
        ms = self._batch.export_states([0])[0]
        self._ms = ms
        h, size = ms.hdr, ms.size
        self.map = GameMap(size)
        self.global_unit_id_count = int(h["unit_id_count"])
        self.global_city_id_count = int(h["city_id_count"])
        self.cities = {}
        slots = []
        for s in range(int(h["n_cities"])):
            city = City(ms.cities[s])
            self.cities[city.id] = city
            slots.append(city)
        for p in range(int(h["n_tiles"])):
            idx = int(ms.tiles[p])
            c = ms.cells[idx]
            cell = self.map.map[idx // size][idx % size]
            city = slots[int(c["city_slot"])]
            cell.city_tile = CityTile(((int(c["flags"]) >> 2) & 3) - 1, cell.pos, city.id, int(c["tile_cd"]),
                                      (int(c["flags"]) >> 4) & 7)
            city.city_cells.append(cell)
        for idx in range(size * size):
            c = ms.cells[idx]
            cell = self.map.map[idx // size][idx % size]
            cell.road = int(c["road_q"]) / 4.0
        for r in range(int(h["n_res"])):
            idx = int(ms.res[r])
            c = ms.cells[idx]
            if int(c["amount"]) > 0:
                cell = self.map.map[idx // size][idx % size]
                cell.resource = Resource(S.RES_NAMES[int(c["flags"]) & 3], int(c["amount"]))
                self.map.resources.append(cell)
                self.map.resources_by_type.setdefault(cell.resource.type, []).append(cell)
        team_states = {}
        stats = {"teamStats": {}}
        for t in (0, 1):
            rp = int(h["research"][t])
            team_states[t] = {"researchPoints": rp, "units": {},
                              "researched": {"wood": True, "coal": rp >= 50, "uranium": rp >= 200}}
            stats["teamStats"][t] = {
                "fuelGenerated": int(h["fuel_generated"][t]),
                "resourcesCollected": {n: int(h["collected"][t][k]) for k, n in enumerate(("wood", "coal", "uranium"))},
                "cityTilesBuilt": int(h["city_tiles_built"][t]), "workersBuilt": int(h["workers_built"][t]),
                "cartsBuilt": int(h["carts_built"][t]), "roadsBuilt": int(h["roads_built_q"][t]) / 4.0,
                "roadsPillaged": 0}
        self._slot_of = {}
        for i in range(int(h["n_units"])):
            u = Unit(ms.units[i])
            team_states[u.team]["units"][u.id] = u
            self.map.map[u.pos.y][u.pos.x].units[u.id] = u
            self._slot_of[(u.team, u.id)] = i
        self.state = {"turn": int(h["turn"]), "teamStates": team_states}
        self.stats = stats

    # ------------------------------------------------------------------ the turn (game.py:390-535)
    def run_turn_with_actions(self, actions):
        ms = self._ms
        size = ms.size
        uw = np.zeros(len(ms.units), dtype=np.uint32)
        tb = np.zeros(len(ms.tiles), dtype=np.uint8)
        tile_pos = {int(ms.tiles[p]): p for p in range(int(ms.hdr["n_tiles"]))}
        for act in actions or []:
            if act is None:
                continue
            if isinstance(act, (A.SpawnAction, A.ResearchAction)):
                idx = act.y * size + act.x if 0 <= act.x < size and 0 <= act.y < size else -1
                p = tile_pos.get(idx)
                owner = self.map.get_cell(act.x, act.y).city_tile.team if p is not None else None
                if p is None or tb[p] != 0 or (not isinstance(act, A.ResearchAction) and owner != act.team):
                    self.dropped_actions += 1
                    continue
                tb[p] = (S.TA_RESEARCH if isinstance(act, A.ResearchAction)
                         else S.TA_BUILD_CART if isinstance(act, A.SpawnCartAction) else S.TA_BUILD_WORKER)
            else:
                slot = self._slot_of.get((act.team, act.unit_id))
                if slot is None or uw[slot] != 0:
                    self.dropped_actions += 1
                    continue
                uw[slot] = act.word()
        import torch
        b = self._batch
        b.unit_actions.copy_(torch.from_numpy(uw.astype(np.int64).astype(np.int32)).view(1, -1))
        b.tile_actions.copy_(torch.from_numpy(tb).view(1, -1))
        b.step()
        self._refresh()
        return bool(self._ms.hdr["done"])

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
        return A.action_from_string(command, team)

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
        grid = []
        for row in self.map.map:
            out = []
            for cell in row:
                d = {}
                if cell.get_road() != 0:
                    d["road"] = cell.get_road()
                if cell.resource:
                    d["type"], d["amount"] = cell.resource.type, cell.resource.amount
                out.append(d)
            grid.append(out)
        return {"turn": self.state["turn"], "globalCityIDCount": self.global_city_id_count,
                "globalunit_idCount": self.global_unit_id_count, "teamStates": teams, "map": grid, "cities": cities}
