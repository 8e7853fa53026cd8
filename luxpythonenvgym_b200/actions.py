"""Action classes with the reference's constructor signatures (luxai2021/game/actions.py:44-438).

They are plain records: validation happens on the device (lux_b200_step, optionally with the
MatchController pre-filter).  Every constructor accepts and ignores extra keyword arguments because
agents build all actions with one common kwargs set (examples/agent_policy.py:444-464).
"""
from . import state as S

MOVE, RESEARCH, BUILD_WORKER, BUILD_CART, BUILD_CITY, TRANSFER, PILLAGE = "m", "r", "bw", "bc", "bcity", "t", "p"
_DIRS = {"c": S.DIR_CENTER, "n": S.DIR_NORTH, "e": S.DIR_EAST, "s": S.DIR_SOUTH, "w": S.DIR_WEST}


def _uid(unit_id):
    return int(str(unit_id).split("_")[1])


class Action:
    def __init__(self, action, team):
        self.action, self.team = action, team

    def to_message(self, game=None):
        raise NotImplementedError


class MoveAction(Action):
    def __init__(self, team, unit_id, direction, **kwarg):
        super().__init__(MOVE, team)
        self.unit_id, self.direction = unit_id, direction

    def to_message(self, game=None):
        return "m {} {}".format(self.unit_id, self.direction)

    def word(self):
        return S.encode_move(_DIRS.get(self.direction, 7))


class SpawnAction(Action):
    def __init__(self, action, team, unit_id, x, y, **kwarg):
        super().__init__(action, team)
        self.unit_id, self.x, self.y = unit_id, x, y


class SpawnCartAction(SpawnAction):
    def __init__(self, team, unit_id, x, y, **kwarg):
        super().__init__(BUILD_CART, team, unit_id, x, y)

    def to_message(self, game=None):
        return "bc {} {}".format(self.x, self.y)


class SpawnWorkerAction(SpawnAction):
    def __init__(self, team, unit_id, x, y, **kwarg):
        super().__init__(BUILD_WORKER, team, unit_id, x, y)

    def to_message(self, game=None):
        return "bw {} {}".format(self.x, self.y)


class SpawnCityAction(Action):
    def __init__(self, team, unit_id, **kwarg):
        super().__init__(BUILD_CITY, team)
        self.unit_id = unit_id

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

    def to_message(self, game=None):
        return "t {} {} {} {}".format(self.source_id, self.destination_id, self.resource_type, self.amount)

    def word(self):
        return S.encode_transfer(_uid(self.destination_id), S.RES_CODES[self.resource_type], self.amount)


class PillageAction(Action):
    def __init__(self, team, unit_id, **kwarg):
        super().__init__(PILLAGE, team)
        self.unit_id = unit_id

    def to_message(self, game=None):
        return "p {}".format(self.unit_id)

    def word(self):
        return S.UA_PILLAGE


class ResearchAction(Action):
    def __init__(self, team, x, y, unit_id=None, **kwarg):
        super().__init__(RESEARCH, team)
        self.x, self.y, self.unit_id = x, y, unit_id

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
