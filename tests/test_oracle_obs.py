"""Observation oracle (oracle/lux_oracle.c lux_oracle_observe) against golden vectors produced by the
reference's examples/agent_policy.py AgentPolicy.get_observation on reference Game objects.

Tolerance: none.  The reference computes float64; the golden file stores both the float64 rows and
their float32 rounding, the oracle computes the same float64 expressions and rounds once, so the
float32 rows must be bit-identical."""
import numpy as np
import pytest

from oracle import coracle
from tests import golden_util as G


@pytest.mark.parametrize("name", G.obs_names())
def test_observation_oracle(name):
    z = G.npz("obs")
    ms = G.get_state(z, name + "/state")
    for team in (0, 1):
        want, codes = z["%s/obs%d" % (name, team)], z["%s/ent%d" % (name, team)]
        got, got_codes = coracle.observe(ms, team)
        assert np.array_equal(codes, got_codes)
        assert got.shape == want.shape
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), np.argwhere(want != got)[:5]
        assert np.array_equal(z["%s/obs64_%d" % (name, team)].astype(np.float32), want)
