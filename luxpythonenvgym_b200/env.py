"""Environment layer over the batched engine.

* `BatchedLuxEnvironment` -- BASELINE.json config 4: thousands of envs advance one TURN per `step`;
  the policy sees one observation row per actionable entity of the learning team (the rows the
  reference would produce one `LuxEnvironment.step` at a time, luxai2021/env/lux_env.py:123-161)
  and returns one action code per row.  Opponent = the reference's idle `Agent()`.
* `LuxEnvironment` -- the reference's single-env gym-style API (`reset() -> obs`,
  `step(action_code) -> (obs, reward, done, {})`, one DECISION per step) implemented on the same
  device path, for code written against luxai2021.env.lux_env.LuxEnvironment.
* `Agent`, `AgentPolicy` -- names kept from luxai2021/env/agent.py and examples/agent_policy.py.
"""
import numpy as np
import torch

from . import mapgen
from .batched import BatchedGame

OBS_DIM = 85
N_ACTIONS = 9


class Agent:
    """luxai2021/env/agent.py:11-86 -- the do-nothing opponent and base class for hooks."""

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
        return "agent"

    def set_team(self, team):
        self.team = team

    def set_controller(self, match_controller):
        self.match_controller = match_controller


class AgentPolicy(Agent):
    """examples/agent_policy.py AgentPolicy: 85-float observations, 9 action codes, its reward.
    Observation, action decoding and reward run on the device (lux_b200_observe /
    lux_b200_encode_actions / lux_b200_reward); this object only carries the spaces."""

    def __init__(self, mode="train", model=None):
        super().__init__()
        self.mode, self.model = mode, model
        self.observation_shape = (OBS_DIM,)
        self.n_actions = N_ACTIONS

    def get_agent_type(self):
        return "learning" if self.mode == "train" else "agent"


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


class LuxEnvironment:
    """Single env, one decision per `step` (luxai2021/env/lux_env.py:73-183)."""

    def __init__(self, configs, learning_agent=None, opponent_agent=None, replay_validate=None,
                 replay_folder=None, replay_prefix="replay", device="cuda:0"):
        self.configs = dict(configs or {})
        self.learning_agent = learning_agent or AgentPolicy("train")
        self.opponent_agent = opponent_agent or Agent()
        self.device = device
        self.current_step = 0
        self._rng = np.random.RandomState(self.configs.get("team_seed", 0))
        self._env = None

    def _make(self):
        seed = self.configs.get("seed")
        if seed is None:
            seed = int(self._rng.randint(0, 2 ** 31 - 1))
        state = mapgen.generate(seed, size=self.configs.get("width"))
        if self._env is None or self._env.size != state.size:
            self._env = BatchedGame(1, state.size, device=self.device)
            self._e_cap = self._env.caps["units"] + self._env.caps["tiles"]
            self._codes = torch.zeros((1, self._e_cap), dtype=torch.int32, device=self.device)
            self._rstate = torch.zeros((1, 4), dtype=torch.int32, device=self.device)
        self._env.load_states([state])
        self._rstate.zero_()

    @property
    def game(self):
        return self._env

    def _observe(self):
        obs, ent, cnt = self._env.observe(self._team, self._e_cap)
        self._n = int(cnt.item())
        self._rows = obs[0, : self._n].cpu().numpy()
        self._cursor = 0

    def reset(self):
        self.current_step = 0
        self._make()
        # the reference flips a coin per reset (match_controller.py:118-122); configs["team"] pins it
        team = int(self.configs["team"]) if self.configs.get("team") is not None else int(self._rng.randint(0, 2))
        self.learning_agent.set_team(team)
        self.opponent_agent.set_team(1 - team)
        self._team = torch.full((1,), team, dtype=torch.uint8, device=self.device)
        self._pending_reward = 0.0
        self._done = False
        self._advance_until_decision()
        return self._rows[self._cursor] if not self._done else None

    def _advance_until_decision(self):
        # run turns until the learning team has an actionable entity (match_controller.py:214-324)
        self._observe()
        while self._n == 0 and not self._done:
            self._play_turn()

    def _play_turn(self):
        e = self._env
        e.encode_actions(e._ent, e._cnt, self._codes)
        e.step(prevalidate=True)
        self._pending_reward += float(e.reward(self._team, self._rstate).item())
        self._done = bool(e.hdr[0, 24].item())
        self._codes.zero_()
        if not self._done:
            self._observe()

    def step(self, action_code):
        self._codes[0, self._cursor] = int(action_code)
        self._cursor += 1
        self.current_step += 1
        reward = 0.0
        if self._cursor >= self._n:
            self._play_turn()
            while self._n == 0 and not self._done:
                self._play_turn()
            reward, self._pending_reward = self._pending_reward, 0.0
        obs = None if self._done else self._rows[self._cursor]
        return obs, reward, self._done, {}
