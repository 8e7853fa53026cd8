"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference engine.

Only `tests/`, `oracle/make_golden.py` and the reference arm of `bench.py` may use this
module; nothing under `luxpythonenvgym_b200/` imports it.  It makes
`/root/reference/luxai2021` importable in this container (no node, no gym, no
stable_baselines3) without touching the reference tree:

* `gym` / `stable_baselines3` stub modules (reference import sites:
  luxai2021/env/agent.py:5, luxai2021/env/lux_env.py:5-7);
* `luxai2021.game.game_map.get_n_values` (reference: luxai2021/env/rng/rng.py:5-16 which
  spawns `node rng.js`) is replaced by an exact restatement of the vendored seedrandom
  ARC4 generator (luxai2021/env/rng/seedrandom.js:1, seeded "gen_<seed>" per
  luxai2021/env/rng/rng.js:4), evaluated lazily.

The reference directory does not exist on the GPU box; `available()` says so and the
callers skip.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LUX_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "luxai2021", "game"))


class SeedRandom:
    """seedrandom.js ARC4 (luxai2021/env/rng/seedrandom.js:1), string seeds < 256 chars.

    mixkey() on an empty key array reduces to key[j] = charCode & 255; KSA over 256
    entries with the key repeated; 256 output bytes discarded; random() builds a double
    from 48 bits and tops it up byte-wise until it has >= 52 significant bits.
    """

    def __init__(self, seed_string):
        key = [ord(c) & 255 for c in seed_string]
        if not key:
            key = [0]
        s = list(range(256))
        j = 0
        for i in range(256):
            t = s[i]
            j = (j + key[i % len(key)] + t) & 255
            s[i] = s[j]
            s[j] = t
        self.s, self.i, self.j = s, 0, 0
        self._g(256)

    def _g(self, count):
        s, i, j, r = self.s, self.i, self.j, 0
        for _ in range(count):
            i = (i + 1) & 255
            t = s[i]
            j = (j + t) & 255
            s[i] = s[j]
            s[j] = t
            r = r * 256 + s[(s[i] + s[j]) & 255]
        self.i, self.j = i, j
        return r

    def random(self):
        n = float(self._g(6))
        d = float(256 ** 6)
        x = 0
        while n < 4503599627370496.0:  # 2**52
            n = (n + x) * 256.0
            d *= 256.0
            x = self._g(1)
        while n >= 9007199254740992.0:  # 2**53
            n /= 2.0
            d /= 2.0
            x >>= 1
        return (n + x) / d


class _LazyValues:
    """Stands in for the list returned by get_n_values(seed, N): indexable doubles."""

    def __init__(self, seed):
        self._rng = SeedRandom("gen_%d" % int(seed))
        self._vals = []

    def __getitem__(self, idx):
        while len(self._vals) <= idx:
            self._vals.append(self._rng.random())
        return self._vals[idx]


def _get_n_values(seed, N=100):
    return _LazyValues(seed)


def _install_stubs():
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:  # gym.Env
            pass

        spaces = types.ModuleType("gym.spaces")

        class Discrete:
            def __init__(self, n):
                self.n = n

        class Box:
            def __init__(self, low=None, high=None, shape=None, dtype=None):
                self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        spaces.Discrete, spaces.Box = Discrete, Box
        gym.Env, gym.spaces = Env, spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    if "stable_baselines3" not in sys.modules:
        sb3 = types.ModuleType("stable_baselines3")
        common = types.ModuleType("stable_baselines3.common")
        callbacks = types.ModuleType("stable_baselines3.common.callbacks")

        class BaseCallback:
            def __init__(self, verbose=0):
                self.verbose = verbose

        callbacks.BaseCallback = BaseCallback
        common.callbacks = callbacks
        sb3.common = common
        sys.modules["stable_baselines3"] = sb3
        sys.modules["stable_baselines3.common"] = common
        sys.modules["stable_baselines3.common.callbacks"] = callbacks


_loaded = None


def load():
    """Import the reference package; returns a namespace with the classes tests need."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import luxai2021.game.game_map as game_map
    game_map.get_n_values = _get_n_values
    from luxai2021.game.game import Game
    from luxai2021.game.constants import Constants, LuxMatchConfigs_Default
    from luxai2021.game import actions
    ns = types.SimpleNamespace(
        Game=Game, Constants=Constants, LuxMatchConfigs_Default=LuxMatchConfigs_Default,
        actions=actions, game_map=game_map)
    _loaded = ns
    return ns


def new_game(seed=None, width=None, height=None, map_type=None):
    """Game(configs) as luxai2021/game/game.py:23-34 builds it."""
    ns = load()
    cfg = dict(ns.LuxMatchConfigs_Default)
    cfg["seed"] = seed
    if width is not None:
        cfg["width"] = width
    if height is not None:
        cfg["height"] = height
    if map_type is not None:
        cfg["mapType"] = map_type
    g = ns.Game(cfg)
    g.log = lambda text: None  # the reference logs to ./log.txt (game.py:636-644); keep the CWD clean
    return g


def game_from_replay(replay, step=0):
    """Bootstrap a Game from a Kaggle replay's update stream.

    Follows luxai2021/game/game.py:159-262 (process_updates) with an empty map of the
    replay's size; id counters are taken from the observation because spawning with an
    explicit id does not advance them (game.py:762-765, :821-824).
    """
    ns = load()
    obs = replay["steps"][step][0]["observation"]
    cfg = dict(ns.LuxMatchConfigs_Default)
    cfg.update(seed=None, mapType="empty", width=obs["width"], height=obs["height"])
    g = ns.Game(cfg)
    g.log = lambda text: None
    g.reset(updates=obs["updates"][2:])
    g.global_unit_id_count = obs["globalUnitIDCount"]
    g.global_city_id_count = obs["globalCityIDCount"]
    g.state["turn"] = obs["step"]
    return g


def agent_policy():
    """The reference's example AgentPolicy (examples/agent_policy.py), train mode, no model."""
    load()
    examples = os.path.join(REFERENCE_ROOT, "examples")
    if examples not in sys.path:
        sys.path.insert(0, examples)
    import agent_policy as ap
    return ap.AgentPolicy(mode="train")


def reference_observations(game, team, agent=None):
    """obs rows + entity codes in MatchController's yield order (match_controller.py:271-289)."""
    import numpy as np
    agent = agent or agent_policy()
    rows, codes = [], []
    new_turn = True
    units = game.state["teamStates"][team]["units"]
    slot = 0
    for t in (0, 1):
        for unit in game.state["teamStates"][t]["units"].values():
            if t == team and unit.can_act():
                rows.append(agent.get_observation(game, unit, None, team, new_turn))
                codes.append(slot)
                new_turn = False
            slot += 1
    pos = 0
    for city in game.cities.values():
        for cell in city.city_cells:
            if city.team == team and cell.city_tile.can_act():
                rows.append(agent.get_observation(game, None, cell.city_tile, team, new_turn))
                codes.append(0x4000 | pos)
                new_turn = False
            pos += 1
    obs = np.array(rows, dtype=np.float64).reshape(len(rows), 85)
    return obs, np.array(codes, dtype=np.int16)
