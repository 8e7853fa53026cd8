"""Action classes with the reference's constructor signatures (luxai2021/game/actions.py:44-438).

The turn itself is validated on the device (lux_b200_step).  `is_valid` / `commit_action_update_stats` are the
MatchController-side pre-filter (actions.py:58-116, :143-192, :256-285, :322-345, :368-385, :412-438) evaluated
on the host `Game` view; they exist for `MatchController.take_action`, which buffers one decision at a time and
needs the answer before the turn runs (the batched environments apply the same filter on the device with
LUX_STEP_PREVALIDATE).  Every constructor accepts and ignores extra keyword arguments because agents build all
actions with one common kwargs set (examples/agent_policy.py:444-464).
"""
from . import state as S

MOVE, RESEARCH, BUILD_WORKER, BUILD_CART, BUILD_CITY, TRANSFER, PILLAGE = "m", "r", "bw", "bc", "bcity", "t", "p"
_DIRS = {"c": S.DIR_CENTER, "n": S.DIR_NORTH, "e": S.DIR_EAST, "s": S.DIR_SOUTH, "w": S.DIR_WEST}
NO_UNIT = 32767          # a destination id no unit carries: the engine's "transfer failed" path (KeyError in the reference)


def _uid(unit_id):
    """Numeric part of "u_N"; anything unparseable maps to NO_UNIT (the reference raises KeyError in
    Worker.turn for such a destination, which game.py:480-485 swallows: no transfer, no cooldown)."""
    try:
        n = int(str(unit_id).split("_")[1])
    except (IndexError, ValueError):
        return NO_UNIT
    return n if 0 <= n < NO_UNIT else NO_UNIT


class Action:
    def __init__(self, action, team):
        self.action, self.team = action, team

    def is_valid(self, game, actions_validated, accumulated_stats=None):
        return True

    def commit_action_update_stats(self, game, accumulated_stats):
        return accumulated_stats

    def to_message(self, game=None):
        raise NotImplementedError


class MoveAction(Action):
    def __init__(self, team, unit_id, direction, **kwarg):
        super().__init__(MOVE, team)
        self.unit_id, self.direction = unit_id, direction

    def is_valid(self, game, actions_validated, accumulated_stats=None):
        """actions.py:58-116.  Of the team's units standing on or next to the target cell, none may already
        hold a buffered move into it; the do-nothing CENTER move is refused whenever a teammate stands on or
        next to the mover (the reference compares the target with the MOVER's own cell there, :111)."""
        if self.unit_id is None or self.team is None or self.direction is None:
            return False
        unit = game.get_unit(self.team, self.unit_id)                 # KeyError for a dead unit, like the reference
        if not unit.can_act():
            return False
        target = unit.pos.translate(self.direction, 1)
        if not game.map.in_map(target):
            return False
        cell = game.map.get_cell_by_pos(target)
        if cell.is_city_tile():
            return cell.city_tile.team == unit.team
        claimed = {}
        for other in actions_validated:
            if other.action == MOVE:
                claimed[other.unit_id] = game.get_unit(other.team, other.unit_id).pos.translate(other.direction, 1)
        stays = target == unit.pos
        for near in game.map.get_adjacent_cells(cell) + [cell]:
            for uid, mate in near.units.items():
                if mate.team != self.team or uid == self.unit_id:
                    continue
                if stays or (uid in claimed and claimed[uid] == target):
                    return False
        return True

    def to_message(self, game=None):
        return "m {} {}".format(self.unit_id, self.direction)

    def word(self):
        return S.encode_move(_DIRS.get(self.direction, 7))


class SpawnAction(Action):
    def __init__(self, action, team, unit_id, x, y, **kwarg):
        super().__init__(action, team)
        self.unit_id, self.x, self.y = unit_id, x, y

    def commit_action_update_stats(self, game, accumulated_stats):
        mine = accumulated_stats[self.team]
        mine["unit_count_offset"] = mine.get("unit_count_offset", 0) + 1
        return accumulated_stats

    def is_valid(self, game, actions_validated, accumulated_stats=None):
        """actions.py:159-194: a tile that can act and a team below its unit cap, counting the spawns already
        approved this turn.  (Ownership of the tile is not looked at here; the engine's validate_command
        drops a spawn on a foreign tile.)"""
        if self.x is None or self.y is None or self.team is None or self.unit_id is not None:
            return False
        if not (0 <= self.x < game.map.width and 0 <= self.y < game.map.height):
            return False
        tile = game.map.get_cell(self.x, self.y).city_tile
        if tile is None or not tile.can_build_unit():
            return False
        offset = (accumulated_stats or {}).get(self.team, {}).get("unit_count_offset", 0)
        return not game.worker_unit_cap_reached(self.team, offset=offset)


class SpawnCartAction(SpawnAction):
    def __init__(self, team, unit_id, x, y, **kwarg):
        super().__init__(BUILD_CART, team, unit_id, x, y)
        self.type = 1

    def to_message(self, game=None):
        return "bc {} {}".format(self.x, self.y)


class SpawnWorkerAction(SpawnAction):
    def __init__(self, team, unit_id, x, y, **kwarg):
        super().__init__(BUILD_WORKER, team, unit_id, x, y)
        self.type = 0

    def to_message(self, game=None):
        return "bw {} {}".format(self.x, self.y)


class SpawnCityAction(Action):
    def __init__(self, team, unit_id, **kwarg):
        super().__init__(BUILD_CITY, team)
        self.unit_id = unit_id

    def is_valid(self, game, actions_validated, accumulated_stats=None):
        """actions.py:256-285"""
        if self.unit_id is None or self.team is None:
            return False
        unit = game.get_unit(self.team, self.unit_id)
        if not unit.can_act() or not unit.can_build(game.map):
            return False
        cell = game.map.get_cell_by_pos(unit.pos)
        return not cell.is_city_tile() and not cell.has_resource()

    def to_message(self, game=None):
        return "bcity {}".format(self.unit_id)

    def word(self):
        return S.UA_BCITY


class TransferAction(Action):
    def __init__(self, team, source_id, destination_id, resource_type, amount, **kwarg):
        super().__init__(TRANSFER, team)
        self.source_id, self.destination_id = source_id, destination_id
        self.resource_type, self.amount = resource_type, amount
        self.unit_id = source_id

    def is_valid(self, game, actions_validated, accumulated_stats=None):
        """actions.py:322-345"""
        if self.source_id is None or self.destination_id is None or self.team is None or self.resource_type is None:
            return False
        if self.source_id == self.destination_id:
            return False
        src = game.get_unit(self.team, self.source_id)
        dst = game.get_unit(self.team, self.destination_id)
        return src.can_act() and src.pos.is_adjacent(dst.pos)

    def to_message(self, game=None):
        return "t {} {} {} {}".format(self.source_id, self.destination_id, self.resource_type, self.amount)

    def word(self):
        res = S.RES_CODES.get(self.resource_type)
        if res is None:
            # an unknown resource name raises KeyError inside Worker.turn in the reference: same "failed" path
            return S.encode_transfer(NO_UNIT, S.RES_WOOD, 0)
        return S.encode_transfer(_uid(self.destination_id), res, self.amount)


class PillageAction(Action):
    def __init__(self, team, unit_id, **kwarg):
        super().__init__(PILLAGE, team)
        self.unit_id = unit_id

    def is_valid(self, game, actions_validated, accumulated_stats=None):
        """actions.py:368-385"""
        if self.unit_id is None or self.team is None:
            return False
        return game.get_unit(self.team, self.unit_id).can_act()

    def to_message(self, game=None):
        return "p {}".format(self.unit_id)

    def word(self):
        return S.UA_PILLAGE


class ResearchAction(Action):
    def __init__(self, team, x, y, unit_id=None, **kwarg):
        super().__init__(RESEARCH, team)
        self.x, self.y, self.unit_id = x, y, unit_id

    def is_valid(self, game, actions_validated, accumulated_stats=None):
        """actions.py:412-438"""
        if self.x is None or self.y is None or self.team is None or self.unit_id is not None:
            return False
        if not (0 <= self.x < game.map.width and 0 <= self.y < game.map.height):
            return False
        tile = game.map.get_cell(self.x, self.y).city_tile
        return tile is not None and tile.can_research()

    def to_message(self, game=None):
        return "r {} {}".format(self.x, self.y)


def action_from_string(command, team):
    """Kaggle command text -> Action (grammar of luxai2021/game/game.py:299-388)."""
    p = command.split(" ")
    op = p[0]
    if op == "m":
        return MoveAction(team, p[1], p[2])
    if op == "bcity":
        return SpawnCityAction(team, p[1])
    if op == "p":
        return PillageAction(team, p[1])
    if op == "t":
        return TransferAction(team, p[1], p[2], p[3], int(p[4]))
    if op == "bw":
        return SpawnWorkerAction(team, None, int(p[1]), int(p[2]))
    if op == "bc":
        return SpawnCartAction(team, None, int(p[1]), int(p[2]))
    if op == "r":
        return ResearchAction(team, int(p[1]), int(p[2]), None)
    raise ValueError("unknown action %r" % command)
