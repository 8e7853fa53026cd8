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


def test_ratio_features_need_no_division():
    """The CUDA observation kernel computes float32(a / c) as float32(a * (1 / c)) (lux_obs_core.cuh lux_ratio32:
    the FP64 division is the expensive part of a row).  Exhaustive over the numerators a row can hold and the six
    divisors of agent_policy.py:188-424 -- with and without the min(v, 1) cap."""
    a = np.arange(-70000, 70000, dtype=np.int64).astype(np.float64)
    for c in (20.0, 500.0, 100.0, 360.0, 30.0, 200.0):
        ref = (a / c).astype(np.float32)
        alt = (a * (1.0 / c)).astype(np.float32)
        assert np.array_equal(ref.view(np.uint32), alt.view(np.uint32)), c
        assert np.array_equal(np.minimum(a / c, 1.0).astype(np.float32), np.minimum(alt, np.float32(1.0))), c
