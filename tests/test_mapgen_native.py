"""Native (C++) host map generator == Python statement == the reference's 100 golden maps."""
import time

import numpy as np

from luxpythonenvgym_b200 import mapgen
from luxpythonenvgym_b200 import state as S
from tests import golden_util as G

CAPS = dict(units=252, cities=256, tiles=512, res=384)


def test_native_matches_reference_golden_maps():
    z = G.npz("maps")
    for seed in [int(s) for s in z["seeds"]]:
        want = G.get_state(z, "%d" % seed)
        got = mapgen.generate_native(seed, caps=CAPS)
        assert got.size == want.size
        assert want.diff(got) == [], seed
        assert np.array_equal(want.res[: int(want.hdr["n_res"])], got.res[: int(got.hdr["n_res"])])


def test_native_batch_equals_python_and_is_fast():
    seeds = np.arange(770000, 770024)
    for size in (12, 16, 24, 32):
        arr = mapgen.generate_batch(seeds, size, threads=4)
        for i, seed in enumerate(seeds[:6]):
            py = mapgen.generate(int(seed), size=size)
            nat = S.MatchState.from_arrays(size, {k: v[i] for k, v in arr.items()})
            assert py.diff(nat) == [], (size, seed)
            assert np.array_equal(py.res, nat.res) and np.array_equal(py.tiles, nat.tiles)
    t0 = time.perf_counter()
    mapgen.generate_batch(np.arange(1000, 1256), 32)
    assert time.perf_counter() - t0 < 20.0
