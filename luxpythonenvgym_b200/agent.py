"""Agent classes with the reference's hook names (luxai2021/env/agent.py:11-287) and the example policy
agent (examples/agent_policy.py:105-575) on top of the device engine.

`Agent` (idle opponent / base class), `AgentFromReplay`, `AgentWithModel`, `AgentFromStdInOut` are host-side
protocol code and follow the reference hook by hook.  `AgentPolicy` keeps the example agent's interface --
85-float observations, 9 action codes, its reward -- but its observation rows and its "transfer to a nearby
cart / worker" target choice come from the CUDA kernels (lux_b200_observe, lux_b200_encode_actions) run on
the match's device state, not from Python loops over the map.
"""
import sys
import time

import numpy as np

from . import actions as A
from . import state as S
from .constants import Constants

OBS_DIM = 85
N_ACTIONS = 9

try:                                     # gym is optional: only the two space descriptions are used
    from gym import spaces               # pragma: no cover
except Exception:                        # noqa: BLE001
    class _Spaces:
        class Discrete:
            def __init__(self, n):
                self.n = n

        class Box:
            def __init__(self, low=None, high=None, shape=None, dtype=None):
                self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    spaces = _Spaces


class Agent:
    """agent.py:11-86 -- the do-nothing opponent and base class of every agent."""

    def __init__(self):
        self.team = None
        self.match_controller = None

    def game_start(self, game):
        pass

    def process_turn(self, game, team):
        return []

    def pre_turn(self, game, is_first_turn=False):
        return

    def post_turn(self, game, actions):
        return False

    def turn_heurstics(self, game, is_first_turn):
        return

    def get_agent_type(self):
        return Constants.AGENT_TYPE.AGENT

    def set_team(self, team):
        self.team = team

    def set_controller(self, match_controller):
        self.match_controller = match_controller


class AgentFromReplay(Agent):
    """agent.py:89-122: replays the commands a Kaggle episode recorded for this agent's team."""

    def __init__(self, replay=None):
        super().__init__()
        self.replay = replay

    def process_turn(self, game, team):
        if self.replay is None:
            return []
        turn = game.state["turn"]
        acts = [game.action_from_string(text, team) for text in self.replay["steps"][turn + 1][team]["action"]]
        return [a for a in acts if a is not None]


class AgentWithModel(Agent):
    """agent.py:126-208: base class of a model-driven agent; "train" mode makes it the learning agent whose
    decisions LuxEnvironment.step supplies, any other mode runs `model.predict` per actionable entity."""

    def __init__(self, mode="train", model=None):
        super().__init__()
        self.action_space = spaces.Discrete(10)
        self.observation_space = spaces.Box(low=0, high=1, shape=(10, 1), dtype=np.float16)
        self.model = model
        self.mode = mode

    def get_agent_type(self):
        return Constants.AGENT_TYPE.LEARNING if self.mode == "train" else Constants.AGENT_TYPE.AGENT

    def get_reward(self, game, is_game_finished, is_new_turn, is_game_error):
        return 0

    def get_observation(self, game, unit, city_tile, team, is_new_turn):
        return np.zeros((10, 1))

    def process_turn(self, game, team):
        started = time.time()
        actions = []
        new_turn = True
        for unit in game.state["teamStates"][team]["units"].values():
            if unit.can_act():
                obs = self.get_observation(game, unit, None, unit.team, new_turn)
                code, _ = self.model.predict(obs, deterministic=False)
                if code is not None:
                    actions.append(self.action_code_to_action(code, game=game, unit=unit, city_tile=None, team=unit.team))
                new_turn = False
        for city in game.cities.values():
            if city.team != team:
                continue
            for cell in city.city_cells:
                tile = cell.city_tile
                if tile.can_act():
                    obs = self.get_observation(game, None, tile, city.team, new_turn)
                    code, _ = self.model.predict(obs, deterministic=False)
                    if code is not None:
                        actions.append(self.action_code_to_action(code, game=game, unit=None, city_tile=tile, team=city.team))
                    new_turn = False
        if time.time() - started > 0.5:
            print("WARNING: Inference took %.3f seconds for computing actions. Limit is 1 second."
                  % (time.time() - started), file=sys.stderr)
        return actions


class AgentFromStdInOut(Agent):
    """agent.py:211-287: the Kaggle stdin/stdout protocol.  Before every turn the engine's update lines are
    read from `stdin` and imported with `game.reset(updates=..., increment_turn=...)` (wire.apply_updates on an
    empty map); after the other agent has decided, its actions are written to `stdout` as command texts and the
    turn is reported as handled, so the local engine does not play it.  `stdin` / `stdout` default to the
    process's own; tests pass file objects."""

    def __init__(self, stdin=None, stdout=None):
        super().__init__()
        self.initialized_player = False
        self.initialized_map = False
        self._in, self._out = stdin, stdout

    def _readline(self):
        if self._in is None:
            return input()
        line = self._in.readline()
        if line == "":
            raise EOFError("end of the update stream")
        return line.rstrip("\n")

    def pre_turn(self, game, is_first_turn=False):
        updates = []
        while True:
            message = self._readline()
            if not self.initialized_player:
                team = int(message)
                self.set_team((team + 1) % 2)
                self.match_controller.set_opponent_team(self, team)
                self.initialized_player = True
            elif not self.initialized_map:
                width, height = message.split(" ")[:2]
                game.configs["width"], game.configs["height"] = int(width), int(height)
                game.configs["mapType"] = Constants.MAP_TYPES.EMPTY
                self.initialized_map = True
            else:
                updates.append(message)
            if message == "D_DONE":
                break
        game.reset(updates=updates, increment_turn=not is_first_turn)

    def post_turn(self, game, actions):
        out = self._out or sys.stdout
        print(",".join(action.to_message(game) for action in actions), file=out)
        print("D_FINISH", file=out)
        out.flush()
        return True


def smart_transfer_to_nearby(game, team, unit_id, unit, target_type_restriction=None, **kwarg):
    """examples/agent_policy.py:24-100: give the largest cargo pile to the best teammate of the wanted type on
    the four neighbouring cells (cart / worker restriction, fullest target that still fits the pile, else the
    emptiest).  The choice is made by the device (csrc/lux_env_core.cuh `env_smart_transfer`, the code behind
    action codes 5 / 6 of lux_b200_encode_actions) on the match's device state; returns a TransferAction, or
    None when there is nothing to give or nobody to give it to."""
    if unit is None:
        unit = game.get_unit(team, unit_id)
    if target_type_restriction not in (Constants.UNIT_TYPES.CART, Constants.UNIT_TYPES.WORKER):
        raise ValueError("target_type_restriction must be UNIT_TYPES.CART or UNIT_TYPES.WORKER")
    slot = game._slot_of[(unit.team, unit.id)]
    word = game._batch.unit_word(0, slot, 5 if target_type_restriction == Constants.UNIT_TYPES.CART else 6)
    d = S.decode_unit_action(word)
    if d["kind"] != S.UA_TRANSFER:
        return None
    return A.TransferAction(unit.team, unit.id, "u_%d" % d["dest"], S.RES_NAMES[d["res"]], d["amount"])


def _unit_table():
    from functools import partial
    return [partial(A.MoveAction, direction="c"), partial(A.MoveAction, direction="n"),
            partial(A.MoveAction, direction="w"), partial(A.MoveAction, direction="s"),
            partial(A.MoveAction, direction="e"),
            partial(smart_transfer_to_nearby, target_type_restriction=Constants.UNIT_TYPES.CART),
            partial(smart_transfer_to_nearby, target_type_restriction=Constants.UNIT_TYPES.WORKER),
            A.SpawnCityAction, A.PillageAction]


def is_stock_tables(agent):
    """True when the agent's action tables are the example's (agent_policy.py:117-132), i.e. what
    lux_b200_encode_actions decodes on the device."""
    from functools import partial
    want_u, want_c = _unit_table(), [A.SpawnWorkerAction, A.SpawnCartAction, A.ResearchAction]
    have_u, have_c = getattr(agent, "actions_units", None), getattr(agent, "actions_cities", None)
    if not isinstance(have_u, list) or not isinstance(have_c, list) or len(have_u) != 9 or have_c != want_c:
        return False
    for a, b in zip(have_u, want_u):
        if isinstance(b, partial):
            if not (isinstance(a, partial) and a.func is b.func and a.args == b.args and a.keywords == b.keywords):
                return False
        elif a is not b:
            return False
    return True


class AgentPolicy(AgentWithModel):
    """examples/agent_policy.py AgentPolicy.  `actions_units` / `actions_cities` are the example's tables
    (:117-132); entries may be replaced by any callable taking the example's keyword arguments (e.g.
    `partial(ActionSequence, actions=[...])`)."""

    def __init__(self, mode="train", model=None):
        super().__init__(mode, model)
        self.actions_units = _unit_table()
        self.actions_cities = [A.SpawnWorkerAction, A.SpawnCartAction, A.ResearchAction]
        self.action_space = spaces.Discrete(max(len(self.actions_units), len(self.actions_cities)))
        self.observation_shape = (OBS_DIM,)
        self.observation_space = spaces.Box(low=0, high=1, shape=self.observation_shape, dtype=np.float16)
        self.n_actions = N_ACTIONS
        self._rows = {}
        self.units_last = self.city_tiles_last = self.fuel_collected_last = 0

    # ---- observation (agent_policy.py:188-424): rows of lux_b200_observe for this turn, looked up by entity
    def get_observation(self, game, unit, city_tile, team, is_new_turn):
        if is_new_turn or getattr(self, "_rows_turn", None) != (id(game), game.state["turn"], team):
            obs, ent, cnt = game._batch.observe(int(team))
            n = int(cnt[0].item())
            rows = obs[0, :n].cpu().numpy()
            codes = ent[0, :n].cpu().numpy()
            self._rows = {int(c): rows[i] for i, c in enumerate(codes)}
            self._rows_turn = (id(game), game.state["turn"], team)
        code = self._entity_code(game, unit, city_tile)
        row = self._rows.get(code)
        if row is None:
            # an entity the engine does not list as actionable (cooldown >= 1) but the controller was told to
            # treat as such (can_act_override): the row the kernel would write is not available
            raise KeyError("no observation row for entity %r (cooldown >= 1?)" % (unit.id if unit else city_tile.get_tile_id()))
        return row.astype(np.float64)

    @staticmethod
    def _entity_code(game, unit, city_tile):
        if unit is not None:
            return game._slot_of[(unit.team, unit.id)]
        size = game.map.width
        idx = city_tile.pos.y * size + city_tile.pos.x
        ms = game._ms
        for p in range(int(ms.hdr["n_tiles"])):
            if int(ms.tiles[p]) == idx:
                return 0x4000 | p
        raise KeyError("not a city tile: %s" % city_tile.pos)

    # ---- action decoding (agent_policy.py:426-478)
    def action_code_to_action(self, action_code, game, unit=None, city_tile=None, team=None):
        try:
            where = city_tile if city_tile is not None else unit
            x, y = (where.pos.x, where.pos.y) if where is not None else (None, None)
            if city_tile is not None:
                make = self.actions_cities[int(action_code) % len(self.actions_cities)]
            else:
                make = self.actions_units[int(action_code) % len(self.actions_units)]
            return make(game=game, unit_id=unit.id if unit else None, unit=unit,
                        city_id=city_tile.city_id if city_tile else None, citytile=city_tile, team=team, x=x, y=y)
        except Exception as e:                      # not a valid action (the reference prints and returns None)
            print(e)
            return None

    def take_action(self, action_code, game, unit=None, city_tile=None, team=None):
        self.match_controller.take_action(self.action_code_to_action(action_code, game, unit, city_tile, team))

    # ---- reward (agent_policy.py:480-560)
    def game_start(self, game):
        self.units_last = self.city_tiles_last = self.fuel_collected_last = 0
        self._rows, self._rows_turn = {}, None

    def get_reward(self, game, is_game_finished, is_new_turn, is_game_error):
        if is_game_error:
            print("Game failed due to error")
            return -1.0
        if not is_new_turn and not is_game_finished:
            return 0
        units = len(game.state["teamStates"][self.team]["units"])
        tiles = sum(len(c.city_cells) for c in game.cities.values() if c.team == self.team)
        fuel = game.stats["teamStats"][self.team]["fuelGenerated"]
        parts = [(units - self.units_last) * 0.05, (tiles - self.city_tiles_last) * 0.1,
                 (fuel - self.fuel_collected_last) / 20000, tiles if is_game_finished else 0]
        self.units_last, self.city_tiles_last, self.fuel_collected_last = units, tiles, fuel
        reward = 0
        for value in parts:
            reward += value
        return reward
