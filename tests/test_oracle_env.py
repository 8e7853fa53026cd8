"""MatchController / LuxEnvironment flow: the oracle's pre-validation, action
decoding and reward against golden episodes driven through the reference's own LuxEnvironment.step
loop (tests/golden/env.npz, oracle/make_golden.make_env)."""
import pytest

from oracle import coracle
from tests import env_util
from tests import golden_util as G


def _step(ms, ent, codes):
    uw, tb = coracle.encode_actions(ms, ent, codes)
    uw, tb = coracle.prevalidate(ms, uw, tb)
    coracle.step(ms, uw, tb)


@pytest.mark.parametrize("k", env_util.env_ids())
def test_env_flow_oracle(k):
    ms = G.get_state(G.npz("env"), "%d/init" % k)
    env_util.check_episode(k, ms, coracle.observe, _step, coracle.reward)
