"""TEST INFRASTRUCTURE -- scripted agents and episode drivers written against the REFERENCE's agent / controller
interface only (luxai2021/env/agent.py, luxai2021/game/match_controller.py, examples/agent_policy.py), so that
the very same code runs on the reference (oracle/make_golden.py records it into tests/golden/hostapi.npz) and on
this package (tests/test_host_api.py replays it, on the emulated backend here and on the GPU).

`ns` is a namespace with the names an agent author imports: Agent, AgentPolicy, ActionSequence, LuxEnvironment,
AgentFromStdInOut, Constants, LuxMatchConfigs_Default, the Action classes and smart_transfer_to_nearby.
"""
import io
import random
import types
from functools import partial

import numpy as np

from oracle import stream


def ours():
    from luxpythonenvgym_b200 import actions as A
    from luxpythonenvgym_b200 import agent, env, match_controller
    from luxpythonenvgym_b200.constants import Constants, LuxMatchConfigs_Default
    return types.SimpleNamespace(
        Agent=agent.Agent, AgentPolicy=agent.AgentPolicy, AgentWithModel=agent.AgentWithModel,
        AgentFromStdInOut=agent.AgentFromStdInOut, AgentFromReplay=agent.AgentFromReplay,
        ActionSequence=match_controller.ActionSequence, LuxEnvironment=env.LuxEnvironment,
        Constants=Constants, LuxMatchConfigs_Default=LuxMatchConfigs_Default,
        MoveAction=A.MoveAction, TransferAction=A.TransferAction, SpawnCityAction=A.SpawnCityAction,
        PillageAction=A.PillageAction, SpawnWorkerAction=A.SpawnWorkerAction, SpawnCartAction=A.SpawnCartAction,
        ResearchAction=A.ResearchAction, smart_transfer_to_nearby=agent.smart_transfer_to_nearby, name="b200")


def reference():
    from oracle import refshim
    ref = refshim.load()
    refshim.agent_policy()                      # puts examples/ on sys.path
    import agent_policy as ap
    from luxai2021.env.agent import Agent, AgentFromReplay, AgentFromStdInOut, AgentWithModel
    from luxai2021.env.lux_env import LuxEnvironment
    from luxai2021.game.match_controller import ActionSequence
    A = ref.actions
    return types.SimpleNamespace(
        Agent=Agent, AgentPolicy=ap.AgentPolicy, AgentWithModel=AgentWithModel, AgentFromStdInOut=AgentFromStdInOut,
        AgentFromReplay=AgentFromReplay, ActionSequence=ActionSequence, LuxEnvironment=LuxEnvironment,
        Constants=ref.Constants, LuxMatchConfigs_Default=ref.LuxMatchConfigs_Default,
        MoveAction=A.MoveAction, TransferAction=A.TransferAction, SpawnCityAction=A.SpawnCityAction,
        PillageAction=A.PillageAction, SpawnWorkerAction=A.SpawnWorkerAction, SpawnCartAction=A.SpawnCartAction,
        ResearchAction=A.ResearchAction, smart_transfer_to_nearby=ap.smart_transfer_to_nearby, name="reference")


def uid_num(unit_id):
    return int(unit_id.split("_")[1])


# --------------------------------------------------------------------------------------------- agents
def make_opponent(ns, seed, k):
    """An inference agent whose `process_turn` returns one hash-coded decision per actionable entity: plain
    actions, smart transfers, action sequences for units and city tiles, and a transfer without destination.
    It counts every hook the controller calls."""
    C = ns.Constants

    class ScriptedOpponent(ns.Agent):
        def __init__(self):
            super().__init__()
            self.calls = {"game_start": 0, "pre_turn": 0, "turn_heurstics": 0, "process_turn": 0, "post_turn": 0}
            self.first_turn_flags = []

        def game_start(self, game):
            self.calls["game_start"] += 1

        def pre_turn(self, game, is_first_turn=False):
            self.calls["pre_turn"] += 1
            self.first_turn_flags.append(bool(is_first_turn))

        def turn_heurstics(self, game, is_first_turn):
            self.calls["turn_heurstics"] += 1

        def post_turn(self, game, actions):
            self.calls["post_turn"] += 1
            self.last_buffer = len(actions)
            return False

        def process_turn(self, game, team):
            self.calls["process_turn"] += 1
            turn = game.state["turn"]
            size = game.map.width
            out = []
            for unit in list(game.state["teamStates"][team]["units"].values()):
                if not unit.can_act():
                    continue
                code = stream.hash64(seed, k, turn, 5, uid_num(unit.id)) % 14
                kw = dict(game=game, unit_id=unit.id, unit=unit, city_id=None, citytile=None, team=team,
                          x=unit.pos.x, y=unit.pos.y)
                if code < 5:
                    out.append(ns.MoveAction(direction="cnwse"[code], **kw))
                elif code == 5:
                    out.append(ns.smart_transfer_to_nearby(target_type_restriction=C.UNIT_TYPES.CART, **kw))
                elif code == 6:
                    out.append(ns.smart_transfer_to_nearby(target_type_restriction=C.UNIT_TYPES.WORKER, **kw))
                elif code in (7, 8):
                    out.append(ns.SpawnCityAction(**kw))
                elif code == 9:
                    out.append(ns.PillageAction(**kw))
                elif code == 10:
                    out.append(ns.ActionSequence(actions=[partial(ns.MoveAction, direction="s"),
                                                          partial(ns.MoveAction, direction="e"),
                                                          ns.SpawnCityAction], **kw))
                elif code == 11:
                    out.append(ns.TransferAction(team, unit.id, None, "wood", 10))
                elif code == 12:
                    out.append(None)
            for city in list(game.cities.values()):
                if city.team != team:
                    continue
                for cell in city.city_cells:
                    tile = cell.city_tile
                    if not tile.can_act():
                        continue
                    code = stream.hash64(seed, k, turn, 6, cell.pos.y * size + cell.pos.x) % 5
                    kw = dict(game=game, unit_id=None, unit=None, city_id=tile.city_id, citytile=tile, team=team,
                              x=cell.pos.x, y=cell.pos.y)
                    if code == 0:
                        out.append(ns.SpawnWorkerAction(**kw))
                    elif code == 1:
                        out.append(ns.SpawnCartAction(**kw))
                    elif code == 2:
                        out.append(ns.ResearchAction(**kw))
                    elif code == 3:
                        out.append(ns.ActionSequence(actions=[ns.ResearchAction, ns.SpawnWorkerAction], **kw))
            return out

    return ScriptedOpponent()


def make_learner(ns, sequences):
    """AgentPolicy("train"); with `sequences` two table entries become ActionSequences (the doc-string usage of
    match_controller.py:15-57).  Counts the hooks LuxEnvironment calls on the learning agent."""

    class CountingPolicy(ns.AgentPolicy):
        def __init__(self):
            super().__init__(mode="train")
            self.calls = {"get_observation": 0, "get_reward": 0, "take_action": 0, "game_start": 0}
            if sequences:
                self.actions_units[5] = partial(ns.ActionSequence, actions=[
                    partial(ns.MoveAction, direction="n"), partial(ns.MoveAction, direction="n"),
                    partial(ns.MoveAction, direction="w")])
                self.actions_units[8] = partial(ns.ActionSequence, actions=[
                    partial(ns.MoveAction, direction="e"), ns.SpawnCityAction])
                self.actions_cities[1] = partial(ns.ActionSequence, actions=[ns.SpawnWorkerAction, ns.ResearchAction])

        def get_observation(self, game, unit, city_tile, team, is_new_turn):
            self.calls["get_observation"] += 1
            return super().get_observation(game, unit, city_tile, team, is_new_turn)

        def get_reward(self, game, is_game_finished, is_new_turn, is_game_error):
            self.calls["get_reward"] += 1
            return super().get_reward(game, is_game_finished, is_new_turn, is_game_error)

        def take_action(self, action_code, game, unit=None, city_tile=None, team=None):
            self.calls["take_action"] += 1
            return super().take_action(action_code, game, unit=unit, city_tile=city_tile, team=team)

        def game_start(self, game):
            self.calls["game_start"] += 1
            return super().game_start(game)

    return CountingPolicy()


# --------------------------------------------------------------------------------------------- episodes
def entity_key(game, unit, tile):
    """(kind, key) of the canonical action stream for the entity a decision is for."""
    if unit is not None:
        return 2, uid_num(unit.id)
    return 3, tile.pos.y * game.map.width + tile.pos.x


def run_env_episode(ns, k, seed, size, sequences, stream_seed, digest, max_turns=360, obs_turns=30, wrap=None):
    """LuxEnvironment(learning = counting AgentPolicy, opponent = scripted agent), decision by decision with
    action code hash64(stream_seed, k, turn, kind, key) % 9.  `digest(game, over)` is evaluated after every
    turn.  Returns a dict of arrays."""
    random.seed(1000 + k)                              # MatchController.reset draws the teams from `random`
    learner, opponent = make_learner(ns, sequences), make_opponent(ns, stream_seed, k)
    cfg = dict(ns.LuxMatchConfigs_Default)
    cfg.update(seed=seed, width=size, height=size)
    env = ns.LuxEnvironment(configs=cfg, learning_agent=learner, opponent_agent=opponent)
    game = env.game
    if wrap:
        wrap(game)
    digests = []
    orig = game.run_turn_with_actions

    def wrapped(actions):
        over = orig(actions)
        digests.append(digest(game, over))
        return over

    game.run_turn_with_actions = wrapped
    obs = env.reset()
    decisions, rows = [], []
    done = False
    while not done and len(digests) < max_turns:
        unit, tile, team, _ = env.last_observation_object
        turn = game.state["turn"]
        kind, key = entity_key(game, unit, tile)
        code = stream.hash64(stream_seed, k, turn, kind, key) % 9
        if turn < obs_turns:
            rows.append(np.asarray(obs, dtype=np.float64))
        obs, reward, done, _ = env.step(code)
        decisions.append((turn, kind, key, code, float(reward), float(done)))
    return {
        "team": np.array(learner.team),
        "decisions": np.array(decisions, dtype=np.float64).reshape(-1, 6),
        "digests": np.array(digests, dtype=np.uint64),
        "obs": np.array(rows, dtype=np.float32).reshape(-1, 85),
        "learner_calls": np.array([learner.calls[n] for n in ("get_observation", "get_reward", "take_action", "game_start")]),
        "opponent_calls": np.array([opponent.calls[n] for n in ("game_start", "pre_turn", "turn_heurstics", "process_turn", "post_turn")]),
        "first_turn_flags": np.array(opponent.first_turn_flags[:4], dtype=np.uint8),
    }


class HashModel:
    """Stands in for a trained SB3 model: `predict(obs)` is a hash of the float32 observation bytes."""

    def predict(self, obs, deterministic=False):
        import hashlib
        raw = np.ascontiguousarray(np.asarray(obs, dtype=np.float32)).tobytes()
        return int.from_bytes(hashlib.blake2b(raw, digest_size=4).digest(), "little") % 9, None


def kaggle_stream(replay, team, turns):
    """The lines a Kaggle engine sends to the agent playing `team`: its id, the map size, then per turn the
    update lines closed by D_DONE (kits/python/simple/main.py of the Lux design; the reference reads them in
    AgentFromStdInOut.pre_turn, agent.py:224-259)."""
    lines = []
    for t in range(turns):
        ups = replay["steps"][t][0]["observation"]["updates"]
        if t == 0:
            lines.append(str(team))                    # the id line; "<w> <h>" follows as updates[1]
            ups = ups[1:]
        lines.extend(ups)
        if not ups or ups[-1] != "D_DONE":
            lines.append("D_DONE")
    return lines


def run_stdin_episode(ns, replay, team, turns):
    """The Kaggle submission wiring (kaggle_submissions/main.py): our model agent in inference mode against
    AgentFromStdInOut, which imports the engine's update lines every turn and prints our commands.  Returns
    the printed lines."""
    import builtins
    import contextlib
    feed = iter(kaggle_stream(replay, team, turns))
    out = io.StringIO()
    player = ns.AgentPolicy(mode="inference", model=HashModel())
    cfg = dict(ns.LuxMatchConfigs_Default)
    remote = ns.AgentFromStdInOut()
    env = ns.LuxEnvironment(configs=cfg, learning_agent=player, opponent_agent=remote)
    real_input = builtins.input

    def fake_input(*_a):
        try:
            return next(feed)
        except StopIteration:
            raise EOFError
    builtins.input = fake_input
    try:
        with contextlib.redirect_stdout(out):
            try:
                env.run_no_learn()
            except EOFError:
                pass
    finally:
        builtins.input = real_input
    return [line for line in out.getvalue().split("\n")]
