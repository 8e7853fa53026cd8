"""Kaggle wire formats (luxpythonenvgym_b200/wire.py) against the reference's own importer: the golden dense
fixtures are replay observations pushed through Game.process_updates of the unmodified reference
(oracle/make_golden.py make_dense / make_wire); the native importer must produce the same packed bytes."""
import numpy as np
import pytest

from luxpythonenvgym_b200 import state as S
from luxpythonenvgym_b200 import wire
from tests import golden_util as G

CAPS = dict(units=252, cities=255, tiles=512, res=384)


def _names():
    return [str(x) for x in G.npz("wire")["names"]]


@pytest.mark.parametrize("name", _names())
def test_import_matches_reference_process_updates(name):
    zw, zd = G.npz("wire"), G.npz("dense")
    size, step, uid, cid = [int(x) for x in zw[name + "/meta"]]
    lines = str(zw[name + "/updates"]).split("\n")
    got = wire.state_from_updates(size, lines, turn=step, unit_id_count=uid, city_id_count=cid, caps=CAPS)
    want = G.get_state(zd, name + "/init")
    assert want.diff(got) == []
    assert got.digest() == want.digest()


@pytest.mark.parametrize("name", _names()[::4])
def test_export_import_round_trip(name):
    zw = G.npz("wire")
    size, step, uid, cid = [int(x) for x in zw[name + "/meta"]]
    a = wire.state_from_updates(size, str(zw[name + "/updates"]).split("\n"), turn=step, unit_id_count=uid,
                                city_id_count=cid, caps=CAPS)
    lines = wire.updates_from_state(a)
    assert lines[-1] == "D_DONE"
    b = wire.state_from_updates(size, lines, turn=step, unit_id_count=uid, city_id_count=cid, caps=CAPS)
    assert a.diff(b) == []
    # and the true neighbour counts instead of the importer's zeros
    c = wire.state_from_updates(size, lines, caps=CAPS, recount_adjacency=True)
    adj = (c.cells["flags"] >> 4) & 7
    assert int(adj.sum()) == int(c.cities["adjsum"][: int(c.hdr["n_cities"])].sum())
    assert int(adj.sum()) > 0


def test_commands_follow_reference_grammar():
    acts = wire.commands_to_actions(["m u_3 n", "bcity u_4", "t u_1 u_2 wood 50", "bw 3 4", "r 5 6", "p u_9"], team=1)
    assert [a.to_message() for a in acts] == ["m u_3 n", "bcity u_4", "t u_1 u_2 wood 50", "bw 3 4", "r 5 6", "p u_9"]
    assert all(a.team == 1 for a in acts)
    for bad in ("m u_3", "bw 1", "x 1 2", "bcity", "t u_1 u_2 wood"):
        with pytest.raises(Exception):
            wire.commands_to_actions([bad], team=0)
