"""`MatchController`, `ActionSequence`, `GameStepFailedException`
(luxai2021/game/match_controller.py:11-324) over the device-backed `Game`.

The controller owns the per-turn protocol agents are written against:
  pre_turn hooks -> pending action sequences -> turn_heurstics hooks -> for each agent either
  `process_turn` (inference agents) or one yielded decision per actionable unit / city tile (the learning
  agent; units in team dict order, then tiles in cities x city_cells order) -> overrides cleared ->
  post_turn hooks -> `game.run_turn_with_actions(action_buffer)` (the CUDA step) -> optional replay check.
`take_action` filters every decision through `Action.is_valid` against the actions already buffered this
turn and marks the entity as having acted, exactly as the reference does, so an agent sees the same
sequence of actionable entities and the engine receives the same action list.
"""
import random
import sys
import traceback

from .agent import Agent
from .constants import Constants
from .game import CityTile


class GameStepFailedException(Exception):
    pass


class ActionSequence:
    """A queue of action constructors handed out one per turn to one unit or city tile, each time the entity
    can act again (match_controller.py:15-71).  `actions` are callables accepting the keyword arguments the
    agent's action table receives (unit_id, citytile, team, x, y, ...)."""

    def __init__(self, actions, unit_id, citytile, **kwarg):
        self.actions = list(actions)
        self.unit_id = unit_id
        self.citytile = citytile
        self.kwarg = kwarg

    def get_next_action(self, game):
        make = self.actions.pop(0)
        return make(unit_id=self.unit_id, citytile=self.citytile, **self.kwarg)

    def is_done(self):
        return not self.actions


class MatchController:
    def __init__(self, game, agents=(None, None), replay_validate=None):
        self.action_buffer = []
        self.game = game
        self.agents = list(agents)
        self.replay_validate = replay_validate
        if len(self.agents) != 2:
            raise ValueError("Two agents must be specified.")
        self.training_agent_count = 0
        for team, agent in enumerate(self.agents):
            if not isinstance(agent, Agent):
                raise ValueError("All agents must inherit from Agent.")
            if agent.get_agent_type() == Constants.AGENT_TYPE.LEARNING:
                self.training_agent_count += 1
            agent.set_team(team)
            agent.set_controller(self)
        if self.training_agent_count > 1:
            raise ValueError("At most one agent must be trainable.")
        self.reset(reset_game=False)

    def reset(self, reset_game=True, randomize_team_order=True, first_team=None):
        """match_controller.py:113-135.  `first_team` (an extension, for reproducible runs) fixes the team of
        agents[0] instead of the coin flip."""
        if first_team is not None:
            self.agents[0].set_team(int(first_team))
            self.agents[1].set_team(1 - int(first_team))
        elif randomize_team_order:
            first = random.randint(0, 1)
            self.agents[0].set_team(first)
            self.agents[1].set_team(1 - first)
        self.action_sequences = {}
        if reset_game:
            self.game.reset()
        self.action_buffer = []
        self.accumulated_stats = {Constants.TEAM.A: {}, Constants.TEAM.B: {}}
        for agent in self.agents:
            agent.game_start(self.game)

    # ------------------------------------------------------------------ buffering decisions (:137-196)
    def take_action(self, action):
        if action is not None:
            if isinstance(action, ActionSequence):
                sequence = action
                if sequence.is_done():
                    return
                action = sequence.get_next_action(self.game)
                if action is None:
                    return
                if not sequence.is_done():
                    if sequence.unit_id is not None:
                        self.action_sequences[sequence.unit_id] = sequence
                    elif sequence.citytile is not None:
                        self.action_sequences[sequence.citytile] = sequence
            try:
                if action.is_valid(self.game, self.action_buffer, self.accumulated_stats):
                    self.action_buffer.append(action)
                    self.accumulated_stats = action.commit_action_update_stats(self.game, self.accumulated_stats)
            except KeyError:
                print("action failed, probably a dead unit %s: %s" % (action, vars(action)), file=sys.stderr)
        # whoever the decision was for has used its turn
        actor = None
        teams = self.game.state["teamStates"]
        uid = getattr(action, "unit_id", None)
        if uid is not None:
            actor = teams[0]["units"].get(uid) or teams[1]["units"].get(uid)
        elif getattr(action, "x", None) is not None:
            cell = self.game.map.get_cell(action.x, action.y)
            if cell.is_city_tile():
                actor = cell.city_tile
        if actor is not None:
            actor.set_can_act_override(False)

    def take_actions(self, actions):
        for action in actions or []:
            self.take_action(action)

    def log_error(self, text):
        try:
            if text is not None:
                with open("match_errors.txt", "a") as out:
                    out.write(text + "\n")
        except Exception:
            print("Critical error in logging")

    def set_opponent_team(self, agent, team):
        for other in self.agents:
            if other is not agent:
                other.set_team(team)

    # ------------------------------------------------------------------ the match loop (:214-324)
    def _continue_sequences(self):
        teams = self.game.state["teamStates"]
        for key in list(self.action_sequences.keys()):
            sequence = self.action_sequences[key]
            actor = None
            if isinstance(key, CityTile):
                if key.city_id in self.game.cities:
                    actor = key
            else:
                actor = teams[0]["units"].get(key) or teams[1]["units"].get(key)
            if actor is None:
                del self.action_sequences[key]          # the entity is gone
            elif actor.can_act():
                self.take_action(sequence.get_next_action(self.game))
                if sequence.is_done():
                    del self.action_sequences[key]

    def run_to_next_observation(self):
        """Generator: yields (unit, city_tile, team, is_new_turn) for every decision of the learning agent and
        plays the turns in between; returns when the match is over."""
        game = self.game
        game_over = False
        is_first_turn = True
        while not game_over:
            turn = game.state["turn"]
            for agent in self.agents:
                agent.pre_turn(game, is_first_turn)
            self._continue_sequences()
            for agent in self.agents:
                agent.turn_heurstics(game, is_first_turn)
            for agent in self.agents:
                kind = agent.get_agent_type()
                if kind == Constants.AGENT_TYPE.AGENT:
                    self.take_actions(agent.process_turn(game, agent.team))
                elif kind == Constants.AGENT_TYPE.LEARNING:
                    new_turn = True
                    for unit in list(game.state["teamStates"][agent.team]["units"].values()):
                        if unit.can_act():
                            yield unit, None, unit.team, new_turn
                            new_turn = False
                    for city in list(game.cities.values()):
                        if city.team != agent.team:
                            continue
                        for cell in city.city_cells:
                            tile = cell.city_tile
                            if tile.can_act():
                                yield None, tile, tile.team, new_turn
                                new_turn = False
            for t in (0, 1):
                for unit in game.state["teamStates"][t]["units"].values():
                    unit.set_can_act_override(None)
            for city in game.cities.values():
                for cell in city.city_cells:
                    cell.city_tile.set_can_act_override(None)
            is_first_turn = False
            try:
                self.accumulated_stats = {Constants.TEAM.A: {}, Constants.TEAM.B: {}}
                handled = False
                for agent in self.agents:
                    if agent.post_turn(game, self.action_buffer):
                        handled = True
                if not handled:
                    game_over = game.run_turn_with_actions(self.action_buffer)
            except Exception as e:
                self.log_error("ERROR: Critical error occurred in turn simulation.")
                self.log_error(repr(e))
                self.log_error("".join(traceback.format_exception(None, e, e.__traceback__)))
                raise GameStepFailedException("Critical error occurred in turn simulation.")
            self.action_buffer = []
            if self.replay_validate is not None:
                game.process_updates(self.replay_validate["steps"][turn + 1][0]["observation"]["updates"], assign=False)
