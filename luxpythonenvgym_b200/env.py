"""Environment layer over the batched engine.

* `BatchedLuxEnvironment` -- BASELINE.json config 4: thousands of envs advance one TURN per `step`;
  the policy sees one observation row per actionable entity of the learning team (the rows the
  reference would produce one `LuxEnvironment.step` at a time, luxai2021/env/lux_env.py:123-161)
  and returns one action code per row.  Opponent = the reference's idle `Agent()`.
* `LuxEnvironment` -- the reference's single-env gym-style API (luxai2021/env/lux_env.py:73-220):
  `LuxEnvironment(configs, learning_agent, opponent_agent, replay_validate, replay_folder, replay_prefix)`,
  `reset() -> obs`, `step(action_code) -> (obs, reward, done, {})` (one DECISION per step), `run_no_learn()`,
  `.game` (a `Game`), `.match_controller`.  The agents' hooks are called exactly as the reference's
  MatchController calls them; every turn runs on the device.  When the learning agent is the stock
  `AgentPolicy` and the opponent the idle base `Agent` (the PPO example's setup, examples/train.py:62-70) and
  no replay option is set, the same episode is produced without the per-decision host hooks: observation,
  action decoding, pre-validation, turn and reward all stay on the device (`fast path`).
* `Agent`, `AgentPolicy`, ... -- re-exported from .agent (names of luxai2021/env/agent.py, examples/agent_policy.py).
"""
import random

import numpy as np
import torch

from . import mapgen
from .agent import (Agent, AgentFromReplay, AgentFromStdInOut, AgentPolicy, AgentWithModel, OBS_DIM, N_ACTIONS,  # noqa: F401
                    is_stock_tables, smart_transfer_to_nearby)
from .batched import BatchedGame
from .constants import Constants
from .game import Game
from .match_controller import GameStepFailedException, MatchController


class BatchedLuxEnvironment:
    def __init__(self, n_envs, size=32, seed=0, pool_maps=64, device="cuda:0", team_seed=12345, packed=False):
        self.n, self.size, self.device = int(n_envs), int(size), torch.device(device)
        self.packed = bool(packed)
        self.game = BatchedGame(self.n, size, device=device)
        self.pool = BatchedGame(pool_maps, size, device=device)
        self._pool_arrays = mapgen.generate_batch(np.arange(seed, seed + pool_maps), size)
        self.pool.load_arrays(self._pool_arrays)
        self.e_cap = self.game.caps["units"] + self.game.caps["tiles"]
        self.team = torch.zeros(self.n, dtype=torch.uint8, device=self.device)
        self.reward_state = torch.zeros((self.n, 4), dtype=torch.int32, device=self.device)
        self.resets = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._gen = torch.Generator(device=self.device).manual_seed(team_seed)
        self._epoch = 0

    def graphed(self, policy, row_cap=None):
        """The whole turn (policy included) as one replayable CUDA graph; see GraphedRollout.  Call after reset()."""
        return GraphedRollout(self, policy, row_cap=row_cap)

    def _draw_teams(self, mask=None):
        # learning team = coin flip per reset (MatchController.reset, match_controller.py:118-122); drawn on the
        # device so that a step never waits for the host
        flips = torch.randint(0, 2, (self.n,), generator=self._gen, dtype=torch.uint8, device=self.device)
        self.team = flips if mask is None else torch.where(mask, flips, self.team)

    def reset(self):
        """All envs start a fresh match; returns (obs [n, e_cap, 85], ent [n, e_cap], count [n]), or with
        packed=True (obs [R, 85], ent [R] (-1 = unused row), count [n], offsets [n+1])."""
        k = self._pool_arrays["hdr"].shape[0]
        idx = np.arange(self.n) % k
        self.game.load_arrays({name: a[idx] for name, a in self._pool_arrays.items()})
        self.reward_state.zero_()
        self._draw_teams()
        return self._observe()

    def _observe(self):
        if self.packed:
            self._last = self.game.observe_packed(self.team)
        else:
            self._last = self.game.observe(self.team, self.e_cap)
        return self._last

    def step(self, codes):
        """codes: int32 [n, e_cap], one AgentPolicy action code per observed entity row.
        Plays one turn everywhere.  Returns ((obs, ent, count), reward [n] float64, done [n] bool);
        finished envs are reset in place (their returned observation is the new match's first one)."""
        g = self.game
        if self.packed:
            _, ent, cnt, offsets = self._last
            g.encode_actions(ent, cnt, codes, offsets=offsets)
        else:
            g.encode_actions(g._ent, g._cnt, codes)
        g.step(prevalidate=True)
        reward = g.reward(self.team, self.reward_state)
        done = g.hdr[:, 24] != 0
        # no host decision here: the reset kernel skips live matches, the other two are masked updates
        g.reset_done(self.pool, self._epoch, self.resets)
        self.reward_state.masked_fill_(done[:, None], 0)
        self._draw_teams(done)
        self._epoch += 1
        return self._observe(), reward, done


class GraphedRollout:
    """One environment turn of a `BatchedLuxEnvironment` as ONE CUDA graph: policy on the current observation
    rows -> lux_b200_encode_actions -> lux_b200_step(prevalidate) -> lux_b200_reward -> masked resets and team
    draws -> lux_b200_observe for the next turn.  Nothing returns to the host inside a turn; the kernels of the
    chain are programmatic dependents of one another inside the graph as they are on a stream.

    `policy(obs_rows[row_cap, 85] float32) -> codes[row_cap] int32` must be capturable (static shapes, no host
    synchronisation); it sees every row of the fixed-capacity buffer, rows without an entity are ignored.
    After `turn()`: `.reward` [n] float64 and `.done` [n] bool of the turn just played, `.obs / .ent / .count /
    .offsets` for the next one, `.decisions` (device int64, running total of entity decisions).
    The tail of the turn is lux_b200_env_finish: team draws are a counter-based hash of (env, turn) instead of the
    eager path's torch generator, and the pool entry of a reset is chosen from the device-side turn counter."""

    def __init__(self, env, policy, row_cap=None, warmup=3):
        g = env.game
        self.env, self.policy = env, policy
        self.row_cap = int(row_cap or 4 * env.n)
        dev = env.device
        self.turns = torch.zeros((), dtype=torch.int32, device=dev)
        self.decisions = torch.zeros((), dtype=torch.int64, device=dev)
        self.reward = torch.zeros(env.n, dtype=torch.float64, device=dev)
        self._done8 = torch.zeros(env.n, dtype=torch.uint8, device=dev)
        self.done = self._done8.view(torch.bool)
        self.obs, self.ent, self.count, self.offsets = g.observe_packed(env.team, row_cap=self.row_cap)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):      # first-call allocations (library scratch, torch workspaces) happen here
                self._turn()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._turn()

    def _turn(self):
        env, g = self.env, self.env.game
        codes = self.policy(self.obs)
        self.decisions += self.count.sum()
        g.encode_actions(self.ent, self.count, codes, offsets=self.offsets)
        g.step(prevalidate=True)
        g.env_finish(env.pool, env.team, env.reward_state, self.reward, self._done8, self.turns, env.resets)
        self.turns += 1
        g.observe_packed(env.team, row_cap=self.row_cap)      # writes obs / ent / count / offsets in place

    def turn(self):
        self.graph.replay()
        return self.reward, self.done


class LuxEnvironment:
    """Single env, one decision per `step` (luxai2021/env/lux_env.py:73-220)."""
    metadata = {"render.modes": ["human"]}

    def __init__(self, configs, learning_agent=None, opponent_agent=None, replay_validate=None,
                 replay_folder=None, replay_prefix="replay", device="cuda:0", fast=None):
        learning_agent = learning_agent if learning_agent is not None else AgentPolicy("train")
        opponent_agent = opponent_agent if opponent_agent is not None else Agent()
        self.configs = dict(configs or {})
        self.game = Game(self.configs, device=device)
        self.match_controller = MatchController(self.game, agents=[learning_agent, opponent_agent],
                                                replay_validate=replay_validate)
        self.replay_prefix, self.replay_folder = replay_prefix, replay_folder
        self.action_space = getattr(learning_agent, "action_space", [])
        self.observation_space = getattr(learning_agent, "observation_space", {})
        self.learning_agent, self.opponent_agent = learning_agent, opponent_agent
        self.device = device
        self.current_step = 0
        self.match_generator = None
        self.last_observation_object = None
        stock = (type(learning_agent) is AgentPolicy and learning_agent.get_agent_type() == Constants.AGENT_TYPE.LEARNING
                 and type(opponent_agent) is Agent and replay_validate is None and not replay_folder
                 and is_stock_tables(learning_agent))
        if fast and not stock:
            raise ValueError("the all-device fast path needs the stock AgentPolicy('train') against the idle Agent() "
                             "and no replay options")
        self._fast = stock if fast is None else bool(fast)
        self._rng = random.Random(self.configs.get("team_seed"))

    def set_replay_path(self, replay_folder, replay_prefix):
        self.replay_prefix, self.replay_folder = replay_prefix, replay_folder
        if replay_folder:
            self._fast = False

    def render(self, **kwargs):
        print(self.current_step)
        print(self.game.map.get_map_string())

    # ------------------------------------------------------------------ the reference flow (hooks on the host)
    def _first_team(self):
        # the reference flips a coin per reset (match_controller.py:118-122); configs["team"] pins the learning
        # agent's team, configs["team_seed"] seeds the coin (both are extensions for reproducible tests)
        team = self.configs.get("team")
        if team is None and self.configs.get("team_seed") is not None:
            team = self._rng.randint(0, 1)
        return team

    def reset(self):
        self.current_step = 0
        self.last_observation_object = None
        if self._fast:
            return self._fast_reset()
        self.match_controller.reset(first_team=self._first_team())
        if self.replay_folder:
            self.game.start_replay_logging(stateful=True, replay_folder=self.replay_folder,
                                           replay_filename_prefix=self.replay_prefix)
        self.match_generator = self.match_controller.run_to_next_observation()
        unit, city_tile, team, is_new_turn = next(self.match_generator)
        obs = self.learning_agent.get_observation(self.game, unit, city_tile, team, is_new_turn)
        self.last_observation_object = (unit, city_tile, team, is_new_turn)
        return obs

    def step(self, action_code):
        if self._fast:
            return self._fast_step(action_code)
        unit, city_tile, team, _ = self.last_observation_object
        self.learning_agent.take_action(action_code, self.game, unit=unit, city_tile=city_tile, team=team)
        self.current_step += 1
        is_new_turn, is_game_over, is_game_error = True, False, False
        try:
            unit, city_tile, team, is_new_turn = next(self.match_generator)
            obs = self.learning_agent.get_observation(self.game, unit, city_tile, team, is_new_turn)
            self.last_observation_object = (unit, city_tile, team, is_new_turn)
        except StopIteration:
            is_game_over, obs = True, None
        except GameStepFailedException:
            is_game_over, is_game_error, obs = True, True, None
        reward = self.learning_agent.get_reward(self.game, is_game_over, is_new_turn, is_game_error)
        return obs, reward, is_game_over, {}

    def run_no_learn(self):
        """Plays a whole match between two inference agents (lux_env.py:194-220); True = the engine failed."""
        for agent in self.match_controller.agents:
            assert agent.get_agent_type() == Constants.AGENT_TYPE.AGENT, "Both agents must be in inference mode"
        self.current_step = 0
        self.last_observation_object = None
        self.match_controller.reset(randomize_team_order=False)
        self.match_generator = self.match_controller.run_to_next_observation()
        is_game_error = False
        try:
            next(self.match_generator)
        except StopIteration:
            print("Episode run finished successfully!")
        except GameStepFailedException:
            is_game_error = True
        return is_game_error

    # ------------------------------------------------------------------ the all-device fast path
    # Same episode as the flow above for AgentPolicy("train") vs Agent(): the learning team's actionable
    # entities come from lux_b200_observe in MatchController's yield order, the codes are decoded by
    # lux_b200_encode_actions, Action.is_valid is LUX_STEP_PREVALIDATE, the reward is lux_b200_reward.
    def _fast_reset(self):
        self.match_controller.reset(first_team=self._first_team())
        b = self.game._batch
        self._e_cap = b.caps["units"] + b.caps["tiles"]
        self._codes = torch.zeros((1, self._e_cap), dtype=torch.int32, device=b.device)
        self._rstate = torch.zeros((1, 4), dtype=torch.int32, device=b.device)
        self._team = torch.full((1,), int(self.learning_agent.team), dtype=torch.uint8, device=b.device)
        self._pending_reward = 0.0
        self._done = False
        self._fast_observe()
        while self._n == 0 and not self._done:
            self._fast_turn()
        return self._rows[self._cursor] if not self._done else None

    def _fast_observe(self):
        obs, ent, cnt = self.game._batch.observe(self._team, self._e_cap)
        self._n = int(cnt.item())
        self._rows = obs[0, : self._n].cpu().numpy()
        self._cursor = 0

    def _fast_turn(self):
        b = self.game._batch
        b.encode_actions(b._ent, b._cnt, self._codes)
        b.step(prevalidate=True)
        self.game._mark_stale()
        self._pending_reward += float(b.reward(self._team, self._rstate).item())
        self._done = bool(b.hdr[0, 24].item())
        self._codes.zero_()
        if not self._done:
            self._fast_observe()

    def _fast_step(self, action_code):
        self._codes[0, self._cursor] = int(action_code)
        self._cursor += 1
        self.current_step += 1
        reward = 0.0
        if self._cursor >= self._n:
            self._fast_turn()
            while self._n == 0 and not self._done:
                self._fast_turn()
            reward, self._pending_reward = self._pending_reward, 0.0
        obs = None if self._done else self._rows[self._cursor]
        return obs, reward, self._done, {}
