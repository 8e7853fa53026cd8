"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
Everything stored is output of the reference engine itself (luxai2021.game.game.Game,
imported through oracle/refshim.py); the packed encoding is oracle/flatten.py.

  maps.npz     100 seeded maps (seeds of luxai2021/tests/testmaps.json, test_map.py:107-144):
               initial packed state of Game({seed}) -- pins the host map generator.
  replays.npz  the 7 Kaggle episodes of luxai2021/tests/replays_for_test (test_replay.py):
               initial state, per-turn packed actions, per-turn state digests, full states
               every 40 turns.
  random.npz   seeded random-action matches on 12/16/24/32 maps (canonical stream,
               oracle/stream.py): initial state, per-turn digests, final state.
  dense.npz    replay states imported at turns 80/200/300 (process_updates, game.py:159-262)
  wire.npz     the update lines those imports consumed (input of luxpythonenvgym_b200/wire.py)
               continued for 45 turns with the canonical stream: digests + final state.
  hostapi.npz  the reference's HOST protocol (LuxEnvironment / MatchController / agents / Game list semantics)
               recorded with the scripted agents of tests/scripted.py: decisions, rewards, observations,
               per-turn digests, hook call counts; the lines AgentFromStdInOut prints; messy raw action lists.
  replay_<id>.json.gz  inputs of the reference's tests/test_replay.py flow, reduced to what the flow reads
               (seed, per-step actions, per-step update lines).
"""
import glob
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import flatten, refshim, stream  # noqa: E402

CAPS = dict(units=252, cities=256, tiles=512, res=384)
OUT = os.path.join(ROOT, "tests", "golden")
SECTIONS = ("hdr", "cells", "units", "cities", "tiles", "res")


def put(store, prefix, ms):
    for name, arr in ms.to_bytes().items():
        store["%s/%s" % (prefix, name)] = arr


def stamp_done(want, game, over):
    if over:
        want.hdr["done"] = 1
        want.hdr["winner"] = flatten.winner_code(game)
    return want


def make_maps(ref):
    tm = json.load(open(os.path.join(refshim.REFERENCE_ROOT, "luxai2021/tests/testmaps.json")))
    store = {}
    seeds = []
    for seed, rows in tm.items():
        g = refshim.new_game(seed=seed)
        # the reference's own assertion (test_map.py:120-143)
        assert len(rows) == g.map.height and len(rows[0]) == g.map.width
        for y, row in enumerate(rows):
            for x, want in enumerate(row):
                cell = g.map.get_cell(x, y)
                assert cell.is_city_tile() == want["citytile"]
                if want["resource"] is None:
                    assert not cell.has_resource()
                else:
                    assert cell.resource.type == want["resource"]["type"]
                    assert cell.resource.amount == want["resource"]["amount"]
        ms = flatten.flatten(g, caps=CAPS)
        seeds.append(int(seed))
        store["%d/size" % int(seed)] = np.array(ms.size)
        put(store, "%d" % int(seed), ms)
    store["seeds"] = np.array(seeds)
    np.savez_compressed(os.path.join(OUT, "maps.npz"), **store)
    print("maps.npz", len(seeds))


def make_replays(ref):
    store = {}
    ids = []
    files = sorted(glob.glob(os.path.join(refshim.REFERENCE_ROOT, "luxai2021/tests/replays_for_test/*[0-9].json")))
    for f in files:
        rid = os.path.basename(f)[:-5]
        rep = json.load(open(f))
        g = refshim.game_from_replay(rep)
        ms = flatten.flatten(g, caps=CAPS)
        store["%s/size" % rid] = np.array(ms.size)
        put(store, "%s/init" % rid, ms)
        ua, ta, digests, inexact = [], [], [], []
        for turn in range(len(rep["steps"]) - 1):
            cmds = [rep["steps"][turn + 1][t]["action"] for t in (0, 1)]
            uw, tb, exact = flatten.commands_to_words(g, cmds, ms)
            if not exact:
                # the packed form drops something of this turn's raw command list (a second command for one
                # entity, a command for a dead unit, ...).  The turn is still pinned: the reference's own
                # assertion below compares the state with what the Kaggle engine produced from the RAW list.
                inexact.append(turn)
            for s in np.nonzero(uw)[0]:
                ua.append((turn, s, uw[s]))
            for s in np.nonzero(tb)[0]:
                ta.append((turn, s, tb[s]))
            over = g.run_turn_with_actions(flatten.words_to_actions(ref, g, uw, tb))
            # the reference's own per-turn assertion (match_controller.py:323-324)
            g.process_updates(rep["steps"][turn + 1][0]["observation"]["updates"], assign=False)
            ms = stamp_done(flatten.flatten(g, caps=CAPS), g, over)
            digests.append(ms.digest())
            if (turn + 1) % 40 == 0 or over:
                put(store, "%s/t%d" % (rid, turn + 1), ms)
            if over:
                break
        store["%s/unit_actions" % rid] = np.array(ua, dtype=np.uint32).reshape(-1, 3)
        store["%s/tile_actions" % rid] = np.array(ta, dtype=np.uint32).reshape(-1, 3)
        store["%s/digests" % rid] = np.array(digests, dtype=np.uint64)
        store["%s/inexact_turns" % rid] = np.array(inexact, dtype=np.int32)
        ids.append(rid)
        print("replay", rid, ms.size, len(digests), "turns the packed form cannot express exactly:", inexact)
    store["ids"] = np.array(ids)
    np.savez_compressed(os.path.join(OUT, "replays.npz"), **store)


def run_stream(ref, g, ms, seed, match_id, max_turns):
    digests = []
    for _ in range(max_turns):
        uw, tb = stream.gen_actions(ms, seed, match_id)
        over = g.run_turn_with_actions(flatten.words_to_actions(ref, g, uw, tb))
        ms = stamp_done(flatten.flatten(g, caps=CAPS), g, over)
        digests.append(ms.digest())
        if over:
            break
    return ms, np.array(digests, dtype=np.uint64)


def make_random(ref, n=32, stream_seed=777):
    store = {}
    for k in range(n):
        size = [12, 16, 24, 32][k % 4]
        g = refshim.new_game(seed=5000 + k, width=size, height=size)
        ms = flatten.flatten(g, caps=CAPS)
        store["%d/size" % k] = np.array(size)
        put(store, "%d/init" % k, ms)
        ms, digests = run_stream(ref, g, ms, stream_seed, k, 360)
        put(store, "%d/final" % k, ms)
        store["%d/digests" % k] = digests
        print("random", k, size, len(digests), int(ms.hdr["n_units"]))
    store["n"] = np.array(n)
    store["stream_seed"] = np.array(stream_seed)
    np.savez_compressed(os.path.join(OUT, "random.npz"), **store)


def make_dense(ref, stream_seed=31337):
    store = {}
    names = []
    files = sorted(glob.glob(os.path.join(refshim.REFERENCE_ROOT, "luxai2021/tests/replays_for_test/*[0-9].json")))
    for f in files:
        rid = os.path.basename(f)[:-5]
        rep = json.load(open(f))
        for start in (80, 200, 300):
            g = refshim.game_from_replay(rep, step=start)
            ms = flatten.flatten(g, caps=CAPS)
            name = "%s@%d" % (rid, start)
            store["%s/size" % name] = np.array(ms.size)
            put(store, "%s/init" % name, ms)
            ms, digests = run_stream(ref, g, ms, stream_seed, start, 45)
            put(store, "%s/final" % name, ms)
            store["%s/digests" % name] = digests
            names.append(name)
            print("dense", name, len(digests), int(ms.hdr["n_units"]))
    store["names"] = np.array(names)
    store["stream_seed"] = np.array(stream_seed)
    np.savez_compressed(os.path.join(OUT, "dense.npz"), **store)


def make_wire():
    """Update lines of the replay observations the dense fixtures were imported from (the expected packed
    states are dense.npz <name>/init, produced by the reference's own process_updates)."""
    store = {}
    names = []
    files = sorted(glob.glob(os.path.join(refshim.REFERENCE_ROOT, "luxai2021/tests/replays_for_test/*[0-9].json")))
    for f in files:
        rid = os.path.basename(f)[:-5]
        rep = json.load(open(f))
        for start in (80, 200, 300):
            obs = rep["steps"][start][0]["observation"]
            name = "%s@%d" % (rid, start)
            store[name + "/updates"] = np.array("\n".join(obs["updates"][2:]))
            store[name + "/meta"] = np.array([obs["width"], obs["step"], obs["globalUnitIDCount"], obs["globalCityIDCount"]])
            names.append(name)
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "wire.npz"), **store)


def make_obs(ref, stream_seed=4711):
    """obs.npz: AgentPolicy.get_observation (examples/agent_policy.py:188-424) of every actionable
    entity of both teams on replay states (turn 80/200/300, +12 turns of the canonical stream so that
    cooldowns/cargo/roads are mixed) and on seeded random matches at turn 25 and 60."""
    store = {}
    names = []
    agent = refshim.agent_policy()

    def snap(name, g):
        ms = flatten.flatten(g, caps=CAPS)
        store["%s/size" % name] = np.array(ms.size)
        put(store, "%s/state" % name, ms)
        for team in (0, 1):
            obs, codes = refshim.reference_observations(g, team, agent)
            store["%s/obs%d" % (name, team)] = obs.astype(np.float32)
            store["%s/obs64_%d" % (name, team)] = obs
            store["%s/ent%d" % (name, team)] = codes
        names.append(name)
        print("obs", name, [len(store["%s/ent%d" % (name, t)]) for t in (0, 1)])

    files = sorted(glob.glob(os.path.join(refshim.REFERENCE_ROOT, "luxai2021/tests/replays_for_test/*[0-9].json")))
    for f in files:
        rid = os.path.basename(f)[:-5]
        rep = json.load(open(f))
        for start in (80, 200, 300):
            g = refshim.game_from_replay(rep, step=start)
            snap("%s@%d" % (rid, start), g)
            ms = flatten.flatten(g, caps=CAPS)
            for _ in range(12):
                uw, tb = stream.gen_actions(ms, stream_seed, start)
                g.run_turn_with_actions(flatten.words_to_actions(ref, g, uw, tb))
                ms = flatten.flatten(g, caps=CAPS)
            snap("%s@%d+12" % (rid, start), g)
    for k in range(8):
        size = [12, 16, 24, 32][k % 4]
        g = refshim.new_game(seed=5000 + k, width=size, height=size)
        snap("rand%d@0" % k, g)
        ms = flatten.flatten(g, caps=CAPS)
        for turn in range(60):
            uw, tb = stream.gen_actions(ms, stream_seed, k)
            if g.run_turn_with_actions(flatten.words_to_actions(ref, g, uw, tb)):
                break
            ms = flatten.flatten(g, caps=CAPS)
            if turn in (24, 59):
                snap("rand%d@%d" % (k, turn + 1), g)
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "obs.npz"), **store)


def make_env(ref, stream_seed=90210, n=8):
    """env.npz: the reference's LuxEnvironment (luxai2021/env/lux_env.py:73-183) driven decision by
    decision: learning agent = examples/agent_policy.py AgentPolicy("train"), opponent = idle Agent(),
    action code of every decision = hash64(stream_seed, match, turn, kind, key) % 9.  Stored per match:
    initial state, learning team, per decision (turn, entity code, action code, returned reward, done),
    per game turn the state digest, and the observation rows of the first 50 turns."""
    import random
    from luxai2021.env.agent import Agent
    from luxai2021.env.lux_env import LuxEnvironment
    store = {}
    for k in range(n):
        size = [12, 16, 24, 32][k % 4]
        random.seed(4000 + k)
        agent = refshim.agent_policy()
        cfg = dict(ref.LuxMatchConfigs_Default)
        cfg.update(seed=7000 + k, width=size, height=size)
        env = LuxEnvironment(configs=cfg, learning_agent=agent, opponent_agent=Agent())
        game = env.game
        game.log = lambda text: None
        digests = []
        orig = game.run_turn_with_actions

        def wrapped(actions, orig=orig, game=game, digests=digests):
            over = orig(actions)
            digests.append(stamp_done(flatten.flatten(game, caps=CAPS), game, over).digest())
            return over

        game.run_turn_with_actions = wrapped
        obs = env.reset()
        init = flatten.flatten(game, caps=CAPS)
        team = agent.team
        decisions, obs_rows = [], []
        done = False
        while not done:
            unit, tile, t, _ = env.last_observation_object
            assert t == team
            turn = game.state["turn"]
            if unit is not None:
                slot = 0
                for tt in (0, 1):
                    for u in game.state["teamStates"][tt]["units"].values():
                        if u is unit:
                            ent = slot
                        slot += 1
                kind, key = 2, int(unit.id.split("_")[1])
            else:
                pos = 0
                for city in game.cities.values():
                    for cell in city.city_cells:
                        if cell.city_tile is tile:
                            ent = 0x4000 | pos
                        pos += 1
                kind, key = 3, tile.pos.y * size + tile.pos.x
            code = stream.hash64(stream_seed, k, turn, kind, key) % 9
            if turn < 50:
                obs_rows.append(np.asarray(obs, dtype=np.float64))
            obs, reward, done, _ = env.step(code)
            decisions.append((turn, ent, code, reward, float(done)))
        store["%d/size" % k] = np.array(size)
        store["%d/team" % k] = np.array(team)
        put(store, "%d/init" % k, init)
        store["%d/decisions" % k] = np.array(decisions, dtype=np.float64).reshape(-1, 5)
        store["%d/digests" % k] = np.array(digests, dtype=np.uint64)
        store["%d/obs" % k] = np.array(obs_rows, dtype=np.float32).reshape(-1, 85)
        print("env", k, size, "team", team, "turns", len(digests), "decisions", len(decisions))
    # dense episodes: the same decision loop (MatchController.run_to_next_observation +
    # AgentPolicy.take_action / get_reward, i.e. the body of LuxEnvironment.step) started from
    # replay states, so that many units of one team compete for cells (MoveAction.is_valid :87-113)
    from luxai2021.game.match_controller import MatchController
    files = sorted(glob.glob(os.path.join(refshim.REFERENCE_ROOT, "luxai2021/tests/replays_for_test/*[0-9].json")))
    k = n
    for f in files:
        rep = json.load(open(f))
        for start in (120, 250):
            random.seed(5000 + k)
            game = refshim.game_from_replay(rep, step=start)
            size = game.map.width
            agent = refshim.agent_policy()
            mc = MatchController(game, agents=[agent, Agent()])
            team = agent.team
            digests = []
            orig = game.run_turn_with_actions

            def wrapped(actions, orig=orig, game=game, digests=digests):
                over = orig(actions)
                digests.append(stamp_done(flatten.flatten(game, caps=CAPS), game, over).digest())
                return over

            game.run_turn_with_actions = wrapped
            init = flatten.flatten(game, caps=CAPS)
            gen = mc.run_to_next_observation()
            decisions = []
            try:
                unit, tile, t, new_turn = next(gen)
                while len(digests) < 30:
                    turn = game.state["turn"]
                    if unit is not None:
                        slot = 0
                        for tt in (0, 1):
                            for u in game.state["teamStates"][tt]["units"].values():
                                if u is unit:
                                    ent = slot
                                slot += 1
                        kind, key = 2, int(unit.id.split("_")[1])
                    else:
                        pos = 0
                        for city in game.cities.values():
                            for cell in city.city_cells:
                                if cell.city_tile is tile:
                                    ent = 0x4000 | pos
                                pos += 1
                        kind, key = 3, tile.pos.y * size + tile.pos.x
                    code = stream.hash64(stream_seed, k, turn, kind, key) % 9
                    agent.take_action(code, game, unit=unit, city_tile=tile, team=t)
                    done = False
                    try:
                        unit, tile, t, new_turn = next(gen)
                    except StopIteration:
                        done, new_turn = True, True
                    reward = agent.get_reward(game, done, new_turn, False)
                    decisions.append((turn, ent, code, reward, float(done)))
                    if done:
                        break
            except StopIteration:
                pass
            store["%d/size" % k] = np.array(size)
            store["%d/team" % k] = np.array(team)
            put(store, "%d/init" % k, init)
            store["%d/decisions" % k] = np.array(decisions, dtype=np.float64).reshape(-1, 5)
            store["%d/digests" % k] = np.array(digests, dtype=np.uint64)
            store["%d/obs" % k] = np.zeros((0, 85), dtype=np.float32)
            print("env-dense", k, os.path.basename(f), start, "team", team, "turns", len(digests), "decisions", len(decisions))
            k += 1
    store["n"] = np.array(k)
    store["stream_seed"] = np.array(stream_seed)
    np.savez_compressed(os.path.join(OUT, "env.npz"), **store)


SLIM_REPLAYS = ("27075871", "26688997", "26691974")


def slim_replay(rep):
    """What LuxEnvironment(replay_validate=rep) + AgentFromReplay(rep) read of a Kaggle episode."""
    steps = []
    for st in rep["steps"]:
        steps.append([{"action": st[0].get("action"), "observation": {"updates": st[0]["observation"].get("updates")}},
                      {"action": st[1].get("action")}])
    return {"configuration": {"seed": rep["configuration"]["seed"]}, "steps": steps}


def messy_commands(ref, g, ms, seed, k, turn):
    """One turn's RAW command list [(team, text)]: the canonical stream plus what real agents get wrong --
    repeated commands, two different commands for one entity, shuffled order (the spawn cap then binds on other
    tiles), invalid directions, moves off the map, commands for the other team's units, dead destinations,
    spawns off a city tile.  Never emitted: pillage / transfer FROM a missing unit and research on a non-tile
    (they raise out of the reference's turn) and a second command for a unit that also
    holds a move (declared outside the packed engine's domain, luxpythonenvgym_b200/game.py)."""
    import random
    rng = random.Random(seed * 7919 + k * 1009 + turn)
    uw, tb = stream.gen_actions(ms, seed, k)
    acts = flatten.words_to_actions(ref, g, uw, tb)
    cmds = [(a.team, a.to_message(g)) for a in acts]
    movers = set(a.unit_id for a in acts if a.action == "m")
    extra = []
    for a in acts:
        r = rng.random()
        if a.action == "m":
            continue
        if r < 0.12:
            extra.append((a.team, a.to_message(g)))                                  # the same command again
        elif r < 0.2 and a.action in ("bcity", "p", "t"):
            extra.append((a.team, ("p %s" if a.action != "p" else "bcity %s") % (a.source_id if a.action == "t" else a.unit_id)))
        elif r < 0.2 and a.action in ("bw", "bc", "r"):
            extra.append((a.team, ("r %d %d" if a.action != "r" else "bw %d %d") % (a.x, a.y)))
    size = g.map.width
    for team in (0, 1):
        for unit in g.state["teamStates"][team]["units"].values():
            if unit.id in movers or any(c[1].split(" ")[1] == unit.id for c in cmds + extra if c[1][0] in "pt" or c[1].startswith("bcity")):
                continue
            r = rng.random()
            if r < 0.05:
                extra.append((team, "m %s q" % unit.id))                              # no such direction
            elif r < 0.10:
                extra.append((1 - team, "m %s n" % unit.id))                          # not this team's unit
            elif r < 0.15:
                extra.append((team, "t %s u_9999 wood 5" % unit.id))                  # destination does not exist
            elif r < 0.18:
                extra.append((team, "bcity %s" % unit.id))
    if rng.random() < 0.3:
        extra.append((rng.randint(0, 1), "bw %d %d" % (size + 3, 1)))                 # off the map
    if rng.random() < 0.3:
        x, y = rng.randrange(size), rng.randrange(size)
        if not g.map.get_cell(x, y).is_city_tile():
            extra.append((rng.randint(0, 1), "bw %d %d" % (x, y)))                    # not a city tile
    cmds = cmds + extra
    rng.shuffle(cmds)
    return cmds


def make_hostapi(ref, stream_seed=4242):
    import gzip
    import random
    sys.path.insert(0, ROOT)
    from tests import scripted
    store = {}
    ns = scripted.reference()

    def digest(game, over):
        return stamp_done(flatten.flatten(game, caps=CAPS), game, over).digest()

    # (1) scripted LuxEnvironment episodes: custom learning agent (hooks counted, ActionSequence table entries)
    #     against a scripted inference agent (process_turn, sequences, smart transfers)
    specs = [(8000, 12, True), (8001, 16, False), (8002, 12, True), (8003, 16, True),
             (8004, 24, False), (8005, 32, True), (8006, 12, False), (8007, 16, True)]
    for k, (seed, size, sequences) in enumerate(specs):
        rec = scripted.run_env_episode(ns, k, seed, size, sequences, stream_seed, digest,
                                       wrap=lambda game: setattr(game, "log", lambda text: None))
        for name, arr in rec.items():
            store["env/%d/%s" % (k, name)] = arr
        store["env/%d/spec" % k] = np.array([seed, size, int(sequences)])
        print("hostapi env", k, size, "turns", len(rec["digests"]), "decisions", len(rec["decisions"]))
    store["env/n"] = np.array(len(specs))
    store["stream_seed"] = np.array(stream_seed)
    # (2) slim replays + the Kaggle stdin / stdout protocol on one of them
    for rid in SLIM_REPLAYS:
        rep = json.load(open(os.path.join(refshim.REFERENCE_ROOT, "luxai2021/tests/replays_for_test/%s.json" % rid)))
        raw = json.dumps(slim_replay(rep), separators=(",", ":")).encode()
        with open(os.path.join(OUT, "replay_%s.json.gz" % rid), "wb") as f:
            f.write(gzip.compress(raw, 9, mtime=0))
    rep = json.load(open(os.path.join(refshim.REFERENCE_ROOT, "luxai2021/tests/replays_for_test/%s.json" % SLIM_REPLAYS[1])))
    for team in (0, 1):
        lines = scripted.run_stdin_episode(ns, rep, team, 80)
        store["stdin/%d" % team] = np.array("\n".join(lines))
        print("hostapi stdin team", team, len(lines), "lines")
    # (3) raw action lists through Game.run_turn_with_actions
    for k in range(6):
        size = [12, 16, 24, 32][k % 4]
        g = refshim.new_game(seed=8100 + k, width=size, height=size)
        ms = flatten.flatten(g, caps=CAPS)
        turns, digests = [], []
        for turn in range(120):
            cmds = messy_commands(ref, g, ms, stream_seed, k, turn)
            turns.append(cmds)
            acts = [g.action_from_string(text, team) for team, text in cmds]
            over = g.run_turn_with_actions([a for a in acts if a is not None])
            ms = stamp_done(flatten.flatten(g, caps=CAPS), g, over)
            digests.append(ms.digest())
            if over:
                break
        store["lists/%d/commands" % k] = np.array(json.dumps(turns))
        store["lists/%d/digests" % k] = np.array(digests, dtype=np.uint64)
        store["lists/%d/spec" % k] = np.array([8100 + k, size])
        print("hostapi lists", k, size, "turns", len(digests), "commands", sum(len(t) for t in turns))
    store["lists/n"] = np.array(6)
    np.savez_compressed(os.path.join(OUT, "hostapi.npz"), **store)


def main():
    """python -m oracle.make_golden [maps replays random dense obs env wire hostapi]  (default: all)"""
    os.makedirs(OUT, exist_ok=True)
    os.chdir("/tmp")
    ref = refshim.load()
    steps = {"maps": make_maps, "replays": make_replays, "random": make_random, "dense": make_dense,
             "obs": make_obs, "env": make_env, "wire": lambda ref: make_wire(), "hostapi": make_hostapi}
    for name in (sys.argv[1:] or list(steps)):
        steps[name](ref)


if __name__ == "__main__":
    main()
