"""The reference's HOST interface on top of the device engine: `Game` (list semantics, `reset(updates=...)`,
`process_updates`), `MatchController` + `ActionSequence`, `LuxEnvironment` (agent hooks, `run_no_learn`,
`replay_validate`), `AgentFromReplay`, `AgentFromStdInOut`, `AgentPolicy`.

Every case runs twice: `backend == "emu"` in the CPU suite (the adapters drive tests/emu_backend.EmuBatchedGame,
i.e. the kernels' source compiled for the host) and `backend == "cuda"` under `-m gpu` (the real BatchedGame ->
lux_b200_* C ABI).  Expected values come from the unmodified reference: tests/golden/hostapi.npz and
replay_<id>.json.gz (oracle/make_golden.py make_hostapi), replays.npz, wire.npz.
"""
import gzip
import io
import json
import os

import numpy as np
import pytest

from luxpythonenvgym_b200 import state as S
from tests import golden_util as G
from tests import scripted

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", marks=pytest.mark.gpu, id="cuda")]
SLIM = ("27075871", "26688997", "26691974")


@pytest.fixture(params=BACKENDS)
def backend(request, monkeypatch):
    if request.param == "emu":
        from tests.emu_backend import use_emulated_batch
        use_emulated_batch(monkeypatch)
    return request.param


def slim_replay(rid):
    with gzip.open(os.path.join(G.GOLDEN, "replay_%s.json.gz" % rid), "rb") as f:
        return json.loads(f.read().decode())


def semantic(ms):
    """What the reference's objects hold of a packed state, independent of how the state was imported
    (list orders of resources, the raw road value under a city tile and the built-counters differ between
    a generated map and an update-line import)."""
    h = ms.hdr
    n_units, n_cities, n_tiles = int(h["n_units"]), int(h["n_cities"]), int(h["n_tiles"])
    units = sorted((int(u["id"]), int(u["team_type"]), int(u["x"]), int(u["y"]), int(u["cd_q"]), int(u["wood"]),
                    int(u["coal"]), int(u["uranium"])) for u in ms.units[:n_units])
    cities = sorted((int(c["id"]), int(c["team"]), int(c["fuel"]), int(c["ntiles"]), int(c["adjsum"])) for c in ms.cities[:n_cities])
    tiles = [int(x) for x in ms.tiles[:n_tiles]]
    tile_info = [(i, int(ms.cells["tile_cd"][i]), int(ms.cells["flags"][i]) >> 2) for i in tiles]
    is_tile = np.zeros(len(ms.cells), dtype=bool)
    is_tile[tiles] = True
    amount = ms.cells["amount"].astype(np.int64)
    rtype = np.where(amount > 0, ms.cells["flags"] & 3, 0)
    road = np.where(is_tile, 0, ms.cells["road_q"])
    return (int(h["turn"]), [int(x) for x in h["research"]], int(h["unit_id_count"]), int(h["city_id_count"]),
            units, cities, tile_info, amount.tolist(), rtype.tolist(), road.tolist())


# ------------------------------------------------------------------ the reference's tests/test_replay.py flow
@pytest.mark.parametrize("rid", SLIM)
def test_run_replay(backend, rid):
    """luxai2021/tests/test_replay.py:8-37 with this package's classes: two AgentFromReplay agents, the Kaggle
    episode as `replay_validate` (the state is asserted against the engine's update lines after every turn),
    `run_no_learn()`.  On top of the reference's assertion the states every 40 turns equal the fixture states
    the reference engine produced."""
    from luxpythonenvgym_b200.constants import LuxMatchConfigs_Default
    from luxpythonenvgym_b200.env import Agent, AgentFromReplay, LuxEnvironment  # noqa: F401
    rep = slim_replay(rid)
    config = LuxMatchConfigs_Default.copy()
    config["seed"] = rep["configuration"]["seed"]
    opponent = AgentFromReplay(replay=rep)
    agent = AgentFromReplay(replay=rep)
    env = LuxEnvironment(configs=config, learning_agent=agent, opponent_agent=opponent, replay_validate=rep)
    z = G.npz("replays")
    checked = []
    game = env.game
    play = game.run_turn_with_actions

    def checking(actions):
        over = play(actions)
        turn = game.state["turn"]
        if "%s/t%d/hdr" % (rid, turn) in z.files:
            assert semantic(game.packed_state()) == semantic(G.get_state(z, "%s/t%d" % (rid, turn))), turn
            checked.append(turn)
        return over

    game.run_turn_with_actions = checking
    is_game_error = env.run_no_learn()
    assert not is_game_error
    assert game.state["turn"] == 360 and len(checked) >= 9
    assert game._batch.steps == 360 if backend == "emu" else True


def test_replay_validate_catches_a_wrong_state(backend):
    """The per-turn check is live: one unit line of the expected stream moved to another cell fails the run."""
    from luxpythonenvgym_b200.constants import LuxMatchConfigs_Default
    from luxpythonenvgym_b200.env import AgentFromReplay, LuxEnvironment
    rep = slim_replay(SLIM[0])
    ups = rep["steps"][30][0]["observation"]["updates"]
    for i, line in enumerate(ups):
        if line.startswith("u "):
            p = line.split(" ")
            p[4] = str((int(p[4]) + 3) % 12)
            ups[i] = " ".join(p)
            break
    config = dict(LuxMatchConfigs_Default, seed=rep["configuration"]["seed"])
    env = LuxEnvironment(config, AgentFromReplay(rep), AgentFromReplay(rep), replay_validate=rep)
    with pytest.raises(AssertionError):
        env.run_no_learn()
    assert env.game.state["turn"] == 30


def test_inexact_replay_turns_are_recorded():
    """oracle/make_golden.make_replays stores which turns' raw Kaggle command lists the packed fixture form
    cannot express (repeated commands, commands for dead units); the raw lists themselves are covered by
    test_run_replay, which feeds them unfiltered."""
    z = G.npz("replays")
    got = {rid: z[rid + "/inexact_turns"].tolist() for rid in G.replay_ids()}
    assert got["27075871"] == [21, 48, 191, 231, 271, 311] and got["26688997"] == []
    assert sum(len(v) for v in got.values()) == 31


# ------------------------------------------------------------------ custom agents through LuxEnvironment
@pytest.mark.parametrize("k", range(8))
def test_scripted_agents_episode(backend, k):
    """A subclassed learning agent (hooks counted; in half the cases ActionSequence entries in its tables)
    against a scripted inference agent (process_turn with plain actions, smart transfers, sequences for units
    and tiles, pre_turn / turn_heurstics / post_turn counted), decision by decision through
    LuxEnvironment.reset / step: entity order, action codes, rewards, done flags, observation rows (float32
    bits), the state after every turn and the number of times each hook ran equal the reference's run."""
    z = G.npz("hostapi")
    seed, size, sequences = [int(x) for x in z["env/%d/spec" % k]]
    rec = scripted.run_env_episode(scripted.ours(), k, seed, size, bool(sequences), int(z["stream_seed"]),
                                   lambda game, over: game.packed_state().digest())
    for name in ("team", "digests", "learner_calls", "opponent_calls", "first_turn_flags"):
        assert np.array_equal(rec[name], z["env/%d/%s" % (k, name)]), name
    want = z["env/%d/decisions" % k]
    assert rec["decisions"].shape == want.shape
    assert np.array_equal(rec["decisions"][:, :4], want[:, :4]) and np.array_equal(rec["decisions"][:, 5], want[:, 5])
    assert np.abs(rec["decisions"][:, 4] - want[:, 4]).max() < 1e-9
    assert np.array_equal(rec["obs"].view(np.uint32), z["env/%d/obs" % k].view(np.uint32))
    assert rec["learner_calls"][0] == len(want) and rec["opponent_calls"][3] == len(rec["digests"])


def test_custom_observation_and_reward_are_used(backend):
    """An AgentWithModel subclass with its own observation / reward / action decoding: LuxEnvironment returns
    exactly what the hooks return and the opponent's process_turn moves its units."""
    from luxpythonenvgym_b200.actions import MoveAction
    from luxpythonenvgym_b200.env import Agent, AgentWithModel, LuxEnvironment

    class Mine(AgentWithModel):
        def __init__(self):
            super().__init__("train")
            self.seen = []

        def get_observation(self, game, unit, city_tile, team, is_new_turn):
            self.seen.append((game.state["turn"], unit.id if unit else city_tile.get_tile_id(), is_new_turn))
            return np.array([game.state["turn"], len(game.get_teams_units(team))], dtype=np.float64)

        def get_reward(self, game, is_game_finished, is_new_turn, is_game_error):
            return 7.5 if is_new_turn else 0.25

        def take_action(self, action_code, game, unit=None, city_tile=None, team=None):
            if unit is not None:
                self.match_controller.take_action(MoveAction(team, unit.id, "nesw"[action_code % 4]))
            else:
                self.match_controller.take_action(None)

    class Walker(Agent):
        def process_turn(self, game, team):
            return [MoveAction(team, u.id, "e") for u in game.get_teams_units(team).values() if u.can_act()]

    env = LuxEnvironment({"seed": 7001, "width": 12, "height": 12, "team": 0}, Mine(), Walker())
    obs = env.reset()
    assert obs.tolist() == [0.0, 1.0] and env.learning_agent.seen[0][2] is True
    start = {u.id: (u.pos.x, u.pos.y) for u in env.game.get_teams_units(1).values()}
    obs, reward, done, info = env.step(1)                # worker; next decision: the city tile of the same turn
    assert reward == 0.25 and not done and env.learning_agent.seen[1][0] == 0
    obs, reward, done, info = env.step(0)                # tile -> the turn runs
    assert reward == 7.5 and obs[0] == 1.0
    moved = {u.id: (u.pos.x, u.pos.y) for u in env.game.get_teams_units(1).values()}
    assert all(moved[i] == (start[i][0] + 1, start[i][1]) for i in start), (start, moved)
    assert env.match_controller is env.learning_agent.match_controller and env.game is env.match_controller.game


def test_stock_agents_take_the_device_fast_path_and_agree(backend):
    """AgentPolicy('train') vs the idle Agent() is recognised and (on the GPU) served without per-decision host
    hooks; a subclass or a non-idle opponent is not.  Both paths give the same episode (env.npz, recorded from the
    reference's LuxEnvironment)."""
    from luxpythonenvgym_b200.env import Agent, AgentPolicy, LuxEnvironment
    z = G.npz("env")
    k = 0
    size, team = int(z["%d/size" % k]), int(z["%d/team" % k])
    dec, want_obs = z["%d/decisions" % k], z["%d/obs" % k]

    class Sub(AgentPolicy):
        pass

    cfg = {"seed": 7000 + k, "width": size, "height": size, "team": team}
    assert LuxEnvironment(cfg, Sub("train"), Agent())._fast is False
    assert LuxEnvironment(cfg, AgentPolicy("train"), Sub("inference", model=scripted.HashModel()))._fast is False
    stock = LuxEnvironment(cfg, AgentPolicy("train"), Agent())
    assert stock._fast is True
    with pytest.raises(ValueError):
        LuxEnvironment(cfg, Sub("train"), Agent(), fast=True)
    envs = [LuxEnvironment(cfg, AgentPolicy("train"), Agent(), fast=False)]
    if backend == "cuda":
        envs.append(stock)
    for env in envs:
        obs = env.reset()
        for i in range(len(dec)):
            if i < len(want_obs):
                assert np.array_equal(np.asarray(obs, dtype=np.float32).view(np.uint32), want_obs[i].view(np.uint32)), i
            obs, reward, done, _ = env.step(int(dec[i, 2]))
            assert abs(reward - float(dec[i, 3])) < 1e-9 and done == bool(dec[i, 4]), i
        assert done


# ------------------------------------------------------------------ Kaggle stdin / stdout agent
@pytest.mark.parametrize("team", [0, 1])
def test_agent_from_stdin_out(backend, team):
    """The Kaggle submission wiring (kaggle_submissions/main.py): a model agent in inference mode plays
    against AgentFromStdInOut, which imports the engine's update lines before every turn
    (Game.reset(updates=..., increment_turn=...)) and prints the chosen commands.  80 turns of a Kaggle episode
    are fed in; the printed lines equal what the reference's classes print."""
    z = G.npz("hostapi")
    want = str(z["stdin/%d" % team]).split("\n")
    got = scripted.run_stdin_episode(scripted.ours(), slim_replay(SLIM[1]), team, 80)
    assert got == want
    assert sum(1 for line in got if line == "D_FINISH") == 80


def test_agent_from_stdin_out_file_objects(backend):
    from luxpythonenvgym_b200.env import Agent, AgentFromStdInOut, LuxEnvironment
    rep = slim_replay(SLIM[1])
    lines = scripted.kaggle_stream(rep, 1, 3)
    out = io.StringIO()
    remote = AgentFromStdInOut(stdin=io.StringIO("\n".join(lines) + "\n"), stdout=out)
    env = LuxEnvironment({}, Agent(), remote)
    with pytest.raises(EOFError):
        env.run_no_learn()
    assert out.getvalue() == "\nD_FINISH\n" * 3
    assert remote.team == 0 and env.learning_agent.team == 1 and env.game.state["turn"] == 2
    assert env.game.map.width == 16 and len(env.game.get_teams_units(0)) >= 1


# ------------------------------------------------------------------ Game.reset(updates=...) / process_updates
def test_game_reset_imports_updates(backend):
    """Game.reset(updates=...) on an empty map is Game.process_updates(assign=True) (game.py:76-157, :159-262):
    the packed state equals the one the reference built from the same lines (wire.npz / dense.npz), the object
    view shows it, increment_turn counts on, and reset() without updates starts the seeded map again."""
    from luxpythonenvgym_b200.game import Game
    zw, zd = G.npz("wire"), G.npz("dense")
    name = [str(x) for x in zw["names"]][0]
    lines = str(zw[name + "/updates"]).split("\n")
    want = G.get_state(zd, name + "/init")
    game = Game({"mapType": "empty", "width": want.size, "height": want.size})
    assert len(game.map.resources) == 0 and not game.cities
    game.reset(updates=lines)
    got = game.packed_state()
    got.hdr["turn"], got.hdr["unit_id_count"], got.hdr["city_id_count"] = want.hdr["turn"], want.hdr["unit_id_count"], want.hdr["city_id_count"]
    assert want.diff(got) == []
    assert sum(len(c.city_cells) for c in game.cities.values()) == int(want.hdr["n_tiles"])
    assert len(game.get_teams_units(0)) + len(game.get_teams_units(1)) == int(want.hdr["n_units"])
    game.process_updates(lines, assign=False)                                  # the imported state passes its own lines
    assert game.state["turn"] == 0
    game.reset(updates=lines, increment_turn=True)
    assert game.state["turn"] == 1
    seeded = Game({"seed": 10002})
    first = seeded.packed_state()
    seeded.run_turn_with_actions([])
    seeded.reset()
    assert first.diff(seeded.packed_state()) == []
    with pytest.raises(TypeError):
        seeded.reset(state=lines)


# ------------------------------------------------------------------ raw action lists
@pytest.mark.parametrize("k", range(6))
def test_raw_action_lists(backend, k):
    """Game.run_turn_with_actions(list) with what real agents get wrong -- repeated commands, two commands for one
    entity (it idles), shuffled order (the spawn cap binds on other tiles), invalid directions, off-map targets,
    the other team's units, dead transfer destinations, spawns off a tile: the state after every turn equals the
    reference's for the same list."""
    from luxpythonenvgym_b200.game import Game
    z = G.npz("hostapi")
    seed, size = [int(x) for x in z["lists/%d/spec" % k]]
    turns = json.loads(str(z["lists/%d/commands" % k]))
    digests = z["lists/%d/digests" % k]
    game = Game({"seed": seed, "width": size, "height": size})
    for t, cmds in enumerate(turns):
        acts = [game.action_from_string(text, team) for team, text in cmds]
        over = game.run_turn_with_actions([a for a in acts if a is not None])
        assert game.packed_state().digest() == int(digests[t]), (k, t)
        assert over == (t == len(turns) - 1 and len(turns) < 120)
    assert game.dropped_actions > 0


def test_turn_raises_like_the_reference_for_missing_entities(backend):
    from luxpythonenvgym_b200 import actions as A
    from luxpythonenvgym_b200.game import Game
    game = Game({"seed": 10002})
    with pytest.raises(KeyError):
        game.run_turn_with_actions([A.PillageAction(0, "u_77")])
    with pytest.raises(AttributeError):
        game.run_turn_with_actions([A.ResearchAction(0, 0, 0, None)])
    uid = next(iter(game.get_teams_units(0)))
    before = game.packed_state()
    assert game.run_turn_with_actions([A.TransferAction(0, uid, None, "wood", 5)]) is False     # swallowed KeyError
    assert int(game.packed_state().units["cd_q"][0]) == int(before.units["cd_q"][0])


def test_adapters_need_the_device():
    """No CPU engine behind the adapters: without the test double and without a CUDA device they raise."""
    import torch
    from luxpythonenvgym_b200.game import Game
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(RuntimeError):
        Game({"seed": 1})
