"""Replay export: commands rebuilt from packed actions reproduce the match when parsed back, the JSON has
the reference writer's keys, and (with the reference tree present) to_state_object matches field by field."""
import json

import numpy as np
import pytest

from luxpythonenvgym_b200 import replay as R
from luxpythonenvgym_b200 import state as S
from oracle import coracle, flatten, refshim
from tests import golden_util as G


def test_commands_round_trip_through_the_text_form(tmp_path):
    z = G.npz("dense")
    name = G.dense_names()[4]
    seed, start = int(z["stream_seed"]), int(name.split("@")[1])
    a = G.get_state(z, name + "/init")
    b = a.copy()
    rep = R.Replay(seed=123, size=a.size, stateful=True)
    for turn in range(20):
        uw, tb = coracle.gen_actions(a, seed, start)
        before = a.copy()
        coracle.step(a, uw, tb)
        rep.add_turn(before, uw, tb, a)
        # parse the texts back (Kaggle grammar) and replay them on the twin state
        cmds = rep.data["allCommands"][-1]
        by_team = [[c["command"] for c in cmds if c["agentID"] == t] for t in (0, 1)]

        class _G:  # minimal view commands_to_words needs
            pass
        uw2, tb2, exact = flatten.commands_to_words(_G(), by_team, b)
        assert exact
        coracle.step(b, uw2, tb2)
        assert a.diff(b) == [], turn
    path = rep.write(str(tmp_path / "replay.json"))
    data = json.load(open(path))
    assert set(data) >= {"seed", "mapType", "teamDetails", "allCommands", "version", "results", "width", "height", "stateful"}
    assert data["version"] == "3.1.0" and len(data["allCommands"]) == 20 and len(data["stateful"]) == 20


@pytest.mark.reference
def test_state_object_matches_reference_to_state_object():
    files = sorted(__import__("glob").glob(refshim.REFERENCE_ROOT + "/luxai2021/tests/replays_for_test/*[0-9].json"))
    rep = json.load(open(files[0]))
    g = refshim.game_from_replay(rep, step=150)
    ms = flatten.flatten(g, caps=dict(units=252, cities=256, tiles=512, res=384))
    want, got = g.to_state_object(), R.state_object(ms)
    assert got["turn"] == want["turn"] and got["cities"].keys() == want["cities"].keys()
    for cid, c in want["cities"].items():
        assert got["cities"][cid] == c
    for t in (0, 1):
        assert got["teamStates"][t]["units"] == want["teamStates"][t]["units"]
        assert got["teamStates"][t]["researchPoints"] == want["teamStates"][t]["researchPoints"]
    for y, row in enumerate(want["map"]):
        for x, d in enumerate(row):
            e = got["map"][y][x]
            assert e.get("road", 0) == d.get("road", 0)
            if d.get("amount", 0) > 0:
                assert e["type"] == d["type"] and e["amount"] == d["amount"]
