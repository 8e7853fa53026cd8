"""-m gpu: the CUDA turn engine (through the C ABI) against the oracle and the golden vectors.

Bar: bit-exact on every state field, every turn.
"""
import numpy as np
import pytest
import torch

from luxpythonenvgym_b200 import _lib, mapgen
from luxpythonenvgym_b200 import state as S
from tests import golden_util as G
from tests import gpu_util

pytestmark = pytest.mark.gpu

BIG = dict(units=252, cities=255, tiles=512, res=384)


def _game(n, size, caps=None):
    from luxpythonenvgym_b200.batched import BatchedGame
    return BatchedGame(n, size, caps)


def _recap(ms, caps):
    o = S.MatchState(ms.size, caps)
    o.hdr, o.cells = ms.hdr.copy(), ms.cells.copy()
    for name in ("units", "cities", "tiles", "res"):
        dst = getattr(o, name)
        src = getattr(ms, name)
        k = min(len(dst), len(src))
        dst[:k] = src[:k]
    return o


@pytest.mark.parametrize("tma,two_class,overlap,group,fork", [(1, -1, 1, 0, 1), (1, 1, 0, 16, 0), (1, 0, 1, 32, 1), (1, 0, 0, 8, 0),
                                                               (1, 1, 1, 32, 1), (1, 1, 1, 32, 0), (1, 1, 1, 8, 0), (1, 1, 0, 32, 1),
                                                               (0, -1, 1, 0, 1)])
def test_replays_bit_exact(tma, two_class, overlap, group, fork):
    """The 7 Kaggle episodes of the reference's test_replay.py, 360 turns each, through every launch
    path: TMA staging with the automatic / forced / disabled two-class launch (the big class as an independent
    grid on the side stream, fork = 1, or as the small class's programmatic primary), 8 / 16 / 32 lanes per match,
    and plain loads."""
    lib = _lib.load()
    lib.lux_b200_set_option(b"group", group)
    lib.lux_b200_set_option(b"tma", tma)
    lib.lux_b200_set_option(b"two_class", two_class)
    lib.lux_b200_set_option(b"overlap", overlap)
    lib.lux_b200_set_option(b"fork", fork)
    try:
        z = G.npz("replays")
        for rid in G.replay_ids():
            init = G.get_state(z, rid + "/init")
            U, T = G.replay_actions(z, rid, init)
            g = _game(1, init.size, BIG)
            g.load_states([_recap(init, BIG)])
            digests = z[rid + "/digests"]
            ua = torch.from_numpy(U[:, : BIG["units"]].astype(np.int64).astype(np.int32)).cuda()
            ta = torch.from_numpy(T[:, : BIG["tiles"]]).cuda()
            for turn in range(len(digests)):
                g.step(ua[turn].contiguous().view(1, -1), ta[turn].contiguous().view(1, -1))
                if turn % 20 == 19 or turn == len(digests) - 1:
                    ms = g.export_states([0])[0]
                    assert ms.digest() == int(digests[turn]), "replay %s turn %d" % (rid, turn)
            assert g.headers()["done"][0] == 1
    finally:
        lib.lux_b200_set_option(b"group", 0)
        lib.lux_b200_set_option(b"tma", 1)
        lib.lux_b200_set_option(b"two_class", -1)
        lib.lux_b200_set_option(b"overlap", 1)
        lib.lux_b200_set_option(b"fork", 1)


def test_random_and_dense_fixtures():
    """Golden random-stream and dense fixtures: device action generator + device step."""
    for fixture, names, base_of in (
            ("random", [str(k) for k in range(int(G.npz("random")["n"]))], lambda nm: int(nm)),
            ("dense", G.dense_names(), lambda nm: int(nm.split("@")[1]))):
        z = G.npz(fixture)
        seed = int(z["stream_seed"])
        for nm in names:
            init = G.get_state(z, nm + "/init")
            digests = z[nm + "/digests"]
            g = _game(1, init.size, BIG)
            g.load_states([_recap(init, BIG)])
            for turn in range(len(digests)):
                g.gen_actions(seed, base_of(nm))
                g.step()
            ms = g.export_states([0])[0]
            assert ms.digest() == int(digests[-1]), "%s %s" % (fixture, nm)
            assert _recap(G.get_state(z, nm + "/final"), BIG).diff(ms) == []


@pytest.mark.parametrize("size,n,turns", [(12, 4096, 360), (16, 1024, 200), (24, 512, 200), (32, 1024, 360)])
def test_batch_vs_oracle_every_turn(size, n, turns):
    """Config 2 of BASELINE.json (4096 parallel 12x12 matches) and the larger sizes:
    per-turn bit-exact diff of every match against the C oracle."""
    n_maps = 32
    maps = [mapgen.generate(1000 + i, size=size) for i in range(n_maps)]
    g = _game(n, size)
    g.load_states([maps[i % n_maps] for i in range(n)])
    host = gpu_util.HostBatch(g)
    seed = 20211 + size
    for turn in range(turns):
        g.gen_actions(seed, 0)
        host.gen_actions(seed, 0)
        if turn % 50 == 0:   # the device stream equals the C statement on every live entity
            hh = host.hdr()
            live = (hh["done"] == 0)[:, None]          # finished matches are not written (nor read) at all
            lu = (np.arange(host.ua.shape[1])[None, :] < hh["n_units"][:, None]) & live
            lt = (np.arange(host.ta.shape[1])[None, :] < hh["n_tiles"][:, None]) & live
            assert np.array_equal(g.unit_actions.cpu().numpy().view(np.uint32)[lu], host.ua[lu]), "action stream differs"
            assert np.array_equal(g.tile_actions.cpu().numpy()[lt], host.ta[lt])
        g.step()
        host.step()
        bad = gpu_util.diff_batches(host, g)
        assert len(bad) == 0, "turn %d: %d matches differ, first %d: %s" % (
            turn, len(bad), bad[0], gpu_util.explain(host, g, int(bad[0])))
    hh = host.hdr()
    assert (hh["status"] == 0).all()
    assert hh["done"].sum() > 0


def test_dense_pool_vs_oracle():
    """Dense late-game states (replay imports) tiled over a batch with different action streams."""
    z = G.npz("dense")
    names = [nm for nm in G.dense_names() if int(z[nm + "/size"]) == 32]
    caps = S.default_caps(32)
    inits = [_recap(G.get_state(z, nm + "/init"), caps) for nm in names]
    n = 512
    g = _game(n, 32)
    g.load_states([inits[i % len(inits)] for i in range(n)])
    host = gpu_util.HostBatch(g)
    for turn in range(60):
        g.gen_actions(99, 0)
        host.gen_actions(99, 0)
        g.step()
        host.step()
        bad = gpu_util.diff_batches(host, g)
        assert len(bad) == 0, "turn %d: %s" % (turn, gpu_util.explain(host, g, int(bad[0])))
    assert (host.hdr()["status"] == 0).all()


def test_done_matches_are_untouched_and_reset():
    size = 12
    maps = [mapgen.generate(50 + i, size=size) for i in range(8)]
    g = _game(64, size)
    g.load_states([maps[i % 8] for i in range(64)])
    pool = _game(8, size)
    pool.load_states(maps)
    for _ in range(120):
        g.gen_actions(5, 0)
        g.step()
    before = {k: t.clone() for k, t in g.sections().items()}
    done = g.headers()["done"].astype(bool)
    assert done.any() and not done.all()
    g.gen_actions(5, 0)
    g.step()
    for k, t in g.sections().items():
        assert torch.equal(t[torch.from_numpy(done).cuda()], before[k][torch.from_numpy(done).cuda()]), k
    st = g.stats().cpu().numpy()
    assert st[0] + st[1] == 64 and st[1] == done.sum()
    count = torch.zeros(1, dtype=torch.int32, device="cuda")
    g.reset_done(pool, 7, count)
    assert int(count.item()) == int(done.sum())
    hh = g.headers()
    assert (hh["done"] == 0).all() and (hh["turn"][done] == 0).all()


def test_bench_trajectory_at_full_size_equals_the_cpu_arm():
    """BASELINE.json config 5's per-GPU shape at FULL size -- 16384 resident 32x32 matches, seeded action stream,
    finished matches replaced from a map pool -- played for 150 turns exactly as bench.py plays it (step, then the
    fused reset + action-stream pass), against the loop bench.py's CPU arm times (lux_oracle_run_batch: stream,
    step, pool reset; 16 host threads over disjoint blocks of matches): every live byte of every match equal, and
    the same number of resets.  Pins the full-size run to the oracle AND the two bench arms to the same work."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import coracle
    n, size, n_pool, turns, seed = 16384, 32, 256, 150, 20211019
    pool_arr = mapgen.generate_batch(np.arange(770000, 770000 + n_pool), size)
    g = _game(n, size)
    pool = _game(n_pool, size)
    pool.load_arrays(pool_arr)
    g.load_arrays({k: np.ascontiguousarray(v[np.arange(n) % n_pool]) for k, v in pool_arr.items()})
    host = gpu_util.HostBatch(g)
    parr = {k: np.ascontiguousarray(v) for k, v in pool_arr.items()}
    pdesc = coracle.batch_desc(size, g.caps, parr, n_pool)
    resets = torch.zeros(1, dtype=torch.int32, device="cuda")
    g.gen_actions(seed, 0)
    for t in range(turns):
        g.step()
        g.reset_gen(pool, t, seed, 0, reset_count=resets)
    torch.cuda.synchronize()
    threads = 16
    chunks = [(k * n // threads, (k + 1) * n // threads) for k in range(threads)]
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(lambda c: coracle.run_batch(host.desc, pdesc, seed, 0, host.ua, host.ta, c[0], c[1] - c[0], turns, 0), chunks))
    bad = gpu_util.diff_batches(host, g)
    assert len(bad) == 0, "%d matches differ, first %d: %s" % (len(bad), bad[0], gpu_util.explain(host, g, int(bad[0])))
    hh = host.hdr()
    assert (hh["status"] == 0).all() and (hh["done"] == 0).all()
    assert int(resets.item()) > n // 8                   # the pool-reset path carried real traffic
    st = g.stats().cpu().numpy()
    assert int(st[0]) == n and int(st[7]) == int(hh["n_units"].sum()) and int(st[8]) == int(hh["n_tiles"].sum())
    assert sum(r[1] for r in res) > n * turns * 0.95     # live match-turns the CPU arm counted


def test_reset_gen_equals_reset_then_gen():
    """The fused harness pass (lux_b200_reset_gen) against the two separate calls: same states, same action
    tensors, same counters, turn by turn on a batch that keeps finishing matches."""
    size, n = 12, 512
    a, b = _game(n, size), _game(n, size)
    pool = _game(32, size)
    arr = mapgen.generate_batch(np.arange(3000, 3000 + n), size)
    pool.load_arrays(mapgen.generate_batch(np.arange(7000, 7032), size))
    for g in (a, b):
        g.load_arrays(arr)
    ca = torch.zeros(2, dtype=torch.int64, device="cuda")
    cb = torch.zeros(2, dtype=torch.int64, device="cuda")
    ra = torch.zeros(1, dtype=torch.int32, device="cuda")
    rb = torch.zeros(1, dtype=torch.int32, device="cuda")
    a.gen_actions(11, 5, counters=ca)
    b.gen_actions(11, 5, counters=cb)
    for turn in range(150):
        a.step()
        b.step()
        a.reset_done(pool, turn, ra)
        a.gen_actions(11, 5, counters=ca)
        b.reset_gen(pool, turn, 11, 5, reset_count=rb, counters=cb)
        assert torch.equal(a.unit_actions, b.unit_actions) and torch.equal(a.tile_actions, b.tile_actions), turn
    for k in a.sections():
        assert torch.equal(a.sections()[k], b.sections()[k]), k
    assert torch.equal(ca, cb) and torch.equal(ra, rb) and int(ra.item()) > 0


def test_step_host_matches_device_step():
    size = 16
    maps = [mapgen.generate(300 + i, size=size) for i in range(4)]
    a, b = _game(32, size), _game(32, size)
    for g in (a, b):
        g.load_states([maps[i % 4] for i in range(32)])
    ua_h = torch.zeros_like(a.unit_actions, device="cpu").pin_memory()
    ta_h = torch.zeros_like(a.tile_actions, device="cpu").pin_memory()
    hdr_h = torch.zeros_like(a.hdr, device="cpu").pin_memory()
    done_h = torch.zeros(32, dtype=torch.uint8).pin_memory()
    for turn in range(40):
        a.gen_actions(3, 0)
        ua_h.copy_(a.unit_actions)
        ta_h.copy_(a.tile_actions)
        a.step()
        hv = hdr_h.numpy().view(S.header_dtype).reshape(-1)
        used = (0, 0) if turn == 0 else (int(hv["n_units"].max()), int(hv["n_tiles"].max()))
        b.step_host(ua_h, ta_h, hdr_h, done_h, u_used=used[0], t_used=used[1])
        assert torch.equal(hdr_h, a.hdr.cpu())
    for k in a.sections():
        assert torch.equal(a.sections()[k], b.sections()[k]), k
    # the done flags alone (no header buffer): some matches made to finish on this turn
    for g in (a, b):
        g.hdr[::5, 0:4] = torch.tensor([103, 1, 0, 0], dtype=torch.uint8, device=g.device)      # turn 359
    a.gen_actions(3, 0)
    ua_h.copy_(a.unit_actions)
    ta_h.copy_(a.tile_actions)
    a.step()
    done_h.zero_()
    b.step_host(ua_h, ta_h, None, done_h)
    assert np.array_equal(done_h.numpy(), a.headers()["done"]) and int(done_h.sum()) >= 7


def test_step_host_sparse_matches_device_step():
    size = 24
    a, b = _game(48, size), _game(48, size)
    arr = mapgen.generate_batch(np.arange(900, 948), size)
    for g in (a, b):
        g.load_arrays(arr)
    ent_h = torch.zeros((48 * 64 + 8, 2), dtype=torch.int32).pin_memory()
    done_h = torch.zeros(48, dtype=torch.uint8).pin_memory()
    summ_h = torch.zeros(4, dtype=torch.int32).pin_memory()
    for turn in range(60):
        a.unit_actions.zero_()
        a.tile_actions.zero_()
        a.gen_actions(17, 0)
        ent = a.pack_entries()
        k = ent.shape[0]
        ent_h[:k].copy_(ent)
        # entries that name no live entity (slot beyond the live lists, a finished match) must be dropped
        ent_h[k, 0], ent_h[k, 1] = (5 << 10) | 200, 1
        ent_h[k + 1, 0], ent_h[k + 1, 1] = -(1 << 31) + ((7 << 10) | 300), 3
        k += 2
        fin = np.nonzero(a.headers()["done"])[0]
        if len(fin):
            ent_h[k, 0], ent_h[k, 1] = (int(fin[0]) << 10) | 0, 1 | (1 << 3)
            k += 1
        a.step(prevalidate=True)
        b.step_host_sparse(ent_h, k, done_h, summ_h, prevalidate=True)
        # the step consumed every action word it played: the tensors are clean for the next list
        assert not bool(b.unit_actions.any()) and not bool(b.tile_actions.any()), turn
        hh = a.headers()
        assert torch.equal(b.hdr, a.hdr), turn
        assert np.array_equal(done_h.numpy(), hh["done"])
        assert summ_h.tolist() == [int(hh["n_units"].max()), int(hh["n_tiles"].max()), int((hh["done"] == 0).sum()),
                                   int((hh["done"] != 0).sum())]
    for k in a.sections():
        assert torch.equal(a.sections()[k], b.sections()[k]), k


def test_step_host_submit_wait_with_work_queued_behind():
    """The two-call form of the host list path: submit, queue pool resets behind it, wait -- flags and summary
    describe the state right after the step (before the resets), and the batch ends equal to the device path's
    step + reset.  32x32 without pre-validation: the consume-only specialisation of the small / big classes."""
    size, n = 32, 512
    a, b = _game(n, size), _game(n, size)
    arr = mapgen.generate_batch(np.arange(5000, 5000 + n), size)
    pool = _game(16, size)
    pool.load_arrays(mapgen.generate_batch(np.arange(77, 93), size))
    for g in (a, b):
        g.load_arrays(arr)
        g.hdr[::3, 0:4] = torch.tensor([94, 1, 0, 0], dtype=torch.uint8, device=g.device)   # turn 350: these end within ten turns
    lib = _lib.load()
    import ctypes
    assert lib.lux_b200_step_host_wait(ctypes.byref(b._desc), None, None, None) == _lib.LUX_E_ARG     # nothing submitted
    ent_h = torch.zeros((n * 64 + 8, 2), dtype=torch.int32).pin_memory()
    done_h = torch.zeros(n, dtype=torch.uint8).pin_memory()
    summ_h = torch.zeros(4, dtype=torch.int32).pin_memory()
    ra = torch.zeros(1, dtype=torch.int32, device=a.device)
    rb = torch.zeros(1, dtype=torch.int32, device=a.device)
    for turn in range(120):
        a.unit_actions.zero_()
        a.tile_actions.zero_()
        a.gen_actions(23, 0)
        ent = a.pack_entries()
        k = ent.shape[0]
        ent_h[:k].copy_(ent)
        a.step()
        hh = a.headers()
        b.step_host_sparse_submit(ent_h, k)
        b.reset_done(pool, turn, rb)                 # queued behind the step, before the host blocks
        b.step_host_wait(done_h, summ_h)
        assert np.array_equal(done_h.numpy(), hh["done"]), turn
        assert summ_h.tolist() == [int(hh["n_units"].max()), int(hh["n_tiles"].max()), int((hh["done"] == 0).sum()),
                                   int((hh["done"] != 0).sum())], turn
        a.reset_done(pool, turn, ra)
        torch.cuda.synchronize()
        assert torch.equal(b.hdr, a.hdr), turn
    assert int(ra.item()) == int(rb.item()) and int(ra.item()) > 0
    for k in a.sections():
        assert torch.equal(a.sections()[k], b.sections()[k]), k


def test_step_is_deterministic_across_launch_forms():
    """The same trajectory from one snapshot, twice per launch form (races between the classes or between a step
    and the pass queued behind it would show as run-to-run differences) and across forms: big class forked to the
    side stream / as the small class's programmatic primary / one full-capacity launch -- byte-equal batches."""
    lib = _lib.load()
    n, size = 4096, 32
    g = _game(n, size)
    pool = _game(64, size)
    pool.load_arrays(mapgen.generate_batch(np.arange(31000, 31064), size))
    g.load_arrays(mapgen.generate_batch(np.arange(32000, 32000 + n), size))
    resets = torch.zeros(1, dtype=torch.int32, device=g.device)
    g.gen_actions(99, 0)
    for t in range(150):
        g.step()
        g.reset_gen(pool, t, 99, 0, reset_count=resets)
    torch.cuda.synchronize()
    snap = {k: v.clone() for k, v in g.sections().items()}
    ua, ta = g.unit_actions.clone(), g.tile_actions.clone()
    want = None
    try:
        # mid: the big class's slice (0 = 12 KB + oversize launch, -1 = full capacity, 5632: most big matches oversize)
        for fork, two, mid, rep in ((1, -1, 0, 0), (1, -1, 0, 1), (0, -1, 0, 0), (0, -1, 0, 1), (1, 0, 0, 0), (1, -1, -1, 0),
                                    (1, -1, 5632, 0), (1, -1, 5632, 1), (1, -1, 5633, 0), (1, -1, 5633, 1),
                                    (1, -1, 0, 8), (1, -1, 0, 9), (1, 1, 0, 4), (1, 1, 0, 5)):
            lib.lux_b200_set_option(b"fork", fork)
            lib.lux_b200_set_option(b"two_class", two)
            lib.lux_b200_set_option(b"cta_sync", (rep // 2) * 2 if rep >= 4 else 0)      # rep >= 4: rendezvous CTAs of 4 / 8 warps
            lib.lux_b200_set_option(b"mid_slice", mid & ~1)
            lib.lux_b200_set_option(b"over_own", -1 if mid <= 0 else mid & 1)      # odd: oversize launch on its own stream
            for k, v in g.sections().items():
                v.copy_(snap[k])
            g.unit_actions.copy_(ua)
            g.tile_actions.copy_(ta)
            for t in range(60):
                g.step()
                g.reset_gen(pool, 150 + t, 99, 0, reset_count=resets)
            torch.cuda.synchronize()
            got = {k: v.clone() for k, v in g.sections().items()}
            if want is None:
                want = got
            for k in want:
                assert torch.equal(want[k], got[k]), (fork, two, mid, rep, k)
        # play order of the small class (heavy matches first through the classification pass's order lists): off,
        # automatic (the default above), eight / three / two forced need classes
        lib.lux_b200_set_option(b"fork", 1)
        lib.lux_b200_set_option(b"two_class", -1)
        lib.lux_b200_set_option(b"mid_slice", 0)
        lib.lux_b200_set_option(b"over_own", -1)
        lib.lux_b200_set_option(b"cta_sync", -1)
        for order, n_ord, step in ((0, 0, 10), (1, 8, 5), (1, 3, 15), (1, 2, 40), (1, 0, 10)):
            for name, v in ((b"order", order), (b"n_ord", n_ord), (b"ord_step", step)):
                assert lib.lux_b200_set_option(name, v) == 0
            for k, v in g.sections().items():
                v.copy_(snap[k])
            g.unit_actions.copy_(ua)
            g.tile_actions.copy_(ta)
            for t in range(60):
                g.step()
                g.reset_gen(pool, 150 + t, 99, 0, reset_count=resets)
            torch.cuda.synchronize()
            for k, v in g.sections().items():
                assert torch.equal(want[k], v), (order, n_ord, step, k)
    finally:
        lib.lux_b200_set_option(b"order", 1)
        lib.lux_b200_set_option(b"n_ord", 0)
        lib.lux_b200_set_option(b"ord_step", 10)
        lib.lux_b200_set_option(b"fork", 1)
        lib.lux_b200_set_option(b"two_class", -1)
        lib.lux_b200_set_option(b"mid_slice", 0)
        lib.lux_b200_set_option(b"over_own", -1)
        lib.lux_b200_set_option(b"cta_sync", -1)


def test_bad_arguments_are_rejected():
    import ctypes
    lib = _lib.load()
    g = _game(4, 12)
    d = _lib.LuxBatch(4, 13, 32, 16, 32, 48, g.hdr.data_ptr(), g.cells.data_ptr(), g.units.data_ptr(),
                      g.cities.data_ptr(), g.tiles.data_ptr(), g.res.data_ptr())
    assert lib.lux_b200_step(ctypes.byref(d), g.unit_actions.data_ptr(), g.tile_actions.data_ptr(), 0, None) == _lib.LUX_E_ARG
    d.size, d.u_cap = 12, 33
    assert lib.lux_b200_step(ctypes.byref(d), g.unit_actions.data_ptr(), g.tile_actions.data_ptr(), 0, None) == _lib.LUX_E_CAPACITY
    d.u_cap = 32
    assert lib.lux_b200_step(ctypes.byref(d), None, g.tile_actions.data_ptr(), 0, None) == _lib.LUX_E_ARG


def test_observations_golden_and_oracle():
    """lux_b200_observe against the reference's AgentPolicy.get_observation (golden rows, bit-exact)
    and against the observation oracle on a large mixed batch."""
    from oracle import coracle
    z = G.npz("obs")
    for size in (12, 16, 24, 32):
        names = [nm for nm in G.obs_names() if int(z[nm + "/size"]) == size]
        states = [_recap(G.get_state(z, nm + "/state"), BIG) for nm in names]
        g = _game(2 * len(states), size, BIG)
        g.load_states(states + states)
        sel = torch.tensor([0] * len(states) + [1] * len(states), dtype=torch.uint8, device="cuda")
        obs, ent, cnt = g.observe(sel)
        obs, ent, cnt = obs.cpu().numpy(), ent.cpu().numpy(), cnt.cpu().numpy()
        for i, nm in enumerate(names):
            for team in (0, 1):
                m = i + team * len(states)
                want, codes = z["%s/obs%d" % (nm, team)], z["%s/ent%d" % (nm, team)]
                assert cnt[m] == len(codes), (nm, team)
                assert np.array_equal(ent[m, : cnt[m]], codes)
                assert np.array_equal(obs[m, : cnt[m]].view(np.uint32), want.view(np.uint32)), (nm, team)
    # large batch after some play, vs the oracle
    maps = [mapgen.generate(2000 + i, size=24) for i in range(16)]
    g = _game(512, 24)
    g.load_states([maps[i % 16] for i in range(512)])
    for turn in range(90):
        g.gen_actions(11, 0)
        g.step()
    host = gpu_util.HostBatch(g)
    sel_np = (np.arange(512) % 2).astype(np.uint8)
    e_cap = g.caps["units"] + g.caps["tiles"]
    obs, ent, cnt = g.observe(torch.from_numpy(sel_np).cuda())
    o_obs, o_ent, o_cnt = coracle.observe_batch(host.desc, sel_np, e_cap)
    cnt = cnt.cpu().numpy()
    assert np.array_equal(cnt, o_cnt)
    live = np.arange(e_cap)[None, :] < cnt[:, None]
    assert np.array_equal(ent.cpu().numpy()[live], o_ent[live])
    assert np.array_equal(obs.cpu().numpy().view(np.uint32)[live], o_obs.view(np.uint32)[live])


def test_env_flow_on_device():
    """observe -> encode_actions -> step(prevalidate) -> reward on the GPU against the golden episodes
    recorded from the reference's LuxEnvironment.step loop (tests/golden/env.npz)."""
    from tests import env_util
    z = G.npz("env")
    for k in env_util.env_ids():
        init = _recap(G.get_state(z, "%d/init" % k), BIG)
        g = _game(1, init.size, BIG)
        g.load_states([init])
        team = int(z["%d/team" % k])
        sel = torch.full((1,), team, dtype=torch.uint8, device="cuda")
        rstate = torch.zeros((1, 4), dtype=torch.int32, device="cuda")

        def observe(ms, t):
            obs, ent, cnt = g.observe(sel)
            c = int(cnt.item())
            return obs[0, :c].cpu().numpy(), ent[0, :c].cpu().numpy()

        def step_turn(ms, ent, codes):
            obs, ent_t, cnt = g._obs, g._ent, g._cnt
            ct = torch.zeros((1, ent_t.shape[1]), dtype=torch.int32, device="cuda")
            ct[0, : len(codes)] = torch.from_numpy(codes).cuda()
            g.encode_actions(ent_t, cnt, ct)
            g.step(prevalidate=True)
            new = g.export_states([0])[0]
            for name in S.MatchState.SECTIONS:
                setattr(ms, name, getattr(new, name))

        def reward(ms, t, last):
            return float(g.reward(sel, rstate).item())

        env_util.check_episode(k, g.export_states([0])[0], observe, step_turn, reward)


def test_game_adapter_replays_reference_episode():
    """The single-match `Game` adapter (reference API names) driven by Kaggle command strings."""
    from luxpythonenvgym_b200.game import Game
    from luxpythonenvgym_b200 import actions as A
    z = G.npz("replays")
    rid = G.replay_ids()[0]
    init = G.get_state(z, rid + "/init")
    game = Game({"seed": 1, "width": init.size})
    game.reset(state=init)
    U, T = G.replay_actions(z, rid, init)
    digests = z[rid + "/digests"]
    for turn in range(120):
        ms = game.packed_state()
        acts = []
        for p in np.nonzero(T[turn])[0]:
            cell = int(ms.tiles[p])
            team = ((int(ms.cells["flags"][cell]) >> 2) & 3) - 1
            cls = {1: A.SpawnWorkerAction, 2: A.SpawnCartAction}.get(int(T[turn][p]))
            x, y = cell % ms.size, cell // ms.size
            acts.append(cls(team, None, x, y) if cls else A.ResearchAction(team, x, y, None))
        for s in np.nonzero(U[turn])[0]:
            d = S.decode_unit_action(U[turn][s])
            team, uid = int(ms.units["team_type"][s]) & 1, "u_%d" % int(ms.units["id"][s])
            if d["kind"] == S.UA_MOVE:
                acts.append(A.MoveAction(team, uid, S.DIR_CHARS[d["dir"]]))
            elif d["kind"] == S.UA_BCITY:
                acts.append(A.SpawnCityAction(team, uid))
            elif d["kind"] == S.UA_PILLAGE:
                acts.append(A.PillageAction(team, uid))
            else:
                acts.append(A.TransferAction(team, uid, "u_%d" % d["dest"], S.RES_NAMES[d["res"]], d["amount"]))
        over = game.run_turn_with_actions(acts)
        assert not over
        assert game.packed_state().digest() == int(digests[turn]), turn
    assert game.state["turn"] == 120 and len(game.cities) > 0
    assert sum(len(c.city_cells) for c in game.cities.values()) == int(game.packed_state().hdr["n_tiles"])
    assert game.map.get_map_string().count("\n") == init.size


def test_lux_environment_single_and_batched():
    from luxpythonenvgym_b200.env import BatchedLuxEnvironment, LuxEnvironment
    env = LuxEnvironment({"seed": 7000, "width": 12, "team_seed": 3})
    obs = env.reset()
    assert obs.shape == (85,)
    total, steps, done = 0.0, 0, False
    while not done and steps < 3000:
        obs, r, done, _ = env.step((steps * 7) % 9)
        total += r
        steps += 1
    assert done and steps > 10 and total != 0.0
    benv = BatchedLuxEnvironment(256, size=16, seed=100, pool_maps=16)
    obs, ent, cnt = benv.reset()
    assert obs.shape == (256, benv.e_cap, 85) and int(cnt.min()) >= 1
    finished = 0
    for t in range(120):
        codes = torch.randint(0, 9, ent.shape, dtype=torch.int32, device="cuda")
        (obs, ent, cnt), reward, done = benv.step(codes)
        finished += int(done.sum())
        assert reward.dtype == torch.float64 and torch.isfinite(reward).all()
    assert finished > 0 and (benv.game.headers()["status"] == 0).all()


def test_lux_environment_matches_reference_decision_by_decision():
    """The single-env `LuxEnvironment` (reset / step(code), one decision per step) against the episodes the
    reference's own LuxEnvironment produced (tests/golden/env.npz fixtures 0-7): every observation row,
    every returned reward and the done flag, decision by decision."""
    from luxpythonenvgym_b200.env import LuxEnvironment
    z = G.npz("env")
    for k in range(8):
        size, team = int(z["%d/size" % k]), int(z["%d/team" % k])
        dec, want_obs = z["%d/decisions" % k], z["%d/obs" % k]
        env = LuxEnvironment({"seed": 7000 + k, "width": size, "team": team})
        obs = env.reset()
        total = 0.0
        for i in range(len(dec)):
            if i < len(want_obs):
                assert np.array_equal(np.asarray(obs, dtype=np.float32).view(np.uint32), want_obs[i].view(np.uint32)), (k, i)
            obs, reward, done, _ = env.step(int(dec[i, 2]))
            total += reward
            assert abs(reward - float(dec[i, 3])) < 1e-9, (k, i, reward, dec[i, 3])
            assert done == bool(dec[i, 4]), (k, i)
        assert done


def test_packed_observations_equal_padded():
    from luxpythonenvgym_b200.env import BatchedLuxEnvironment
    env = BatchedLuxEnvironment(128, size=16, seed=500, pool_maps=16, packed=True)
    obs, ent, cnt, off = env.reset()
    for turn in range(40):
        g = env.game
        pobs, pent, pcnt, poff = g.observe_packed(env.team)
        pobs, pent, pcnt, poff = pobs.clone(), pent.clone(), pcnt.clone(), poff.clone()
        dobs, dent, dcnt = g.observe(env.team)
        assert torch.equal(pcnt, dcnt)
        for m in range(0, 128, 7):
            c, o = int(dcnt[m]), int(poff[m])
            assert torch.equal(pobs[o:o + c], dobs[m, :c]) and torch.equal(pent[o:o + c], dent[m, :c])
            assert (pent[o + c:int(poff[m + 1])] == -1).all()
        codes = torch.randint(0, 9, (obs.shape[0],), dtype=torch.int32, device="cuda")
        (obs, ent, cnt, off), reward, done = env.step(codes)
    assert (env.game.headers()["status"] == 0).all()


def test_env_finish_and_device_offsets_equal_the_separate_calls():
    """lux_b200_env_finish == lux_b200_reward + done flags + lux_b200_reset_done(epoch = turn) + cleared reward state
    (+ the documented team draw), and lux_b200_observe_offsets == the host-side cumulative sum of the packed
    layout, clamped to the row capacity."""
    from luxpythonenvgym_b200.batched import BatchedGame
    n, size = 700, 16                     # not a multiple of the scan block: two blocks, the second partial
    arrays = mapgen.generate_batch(np.arange(600, 600 + 64), size)
    pool = BatchedGame(64, size)
    pool.load_arrays(arrays)
    a, b = BatchedGame(n, size), BatchedGame(n, size)
    for g in (a, b):
        g.load_arrays({k: v[np.arange(n) % 64] for k, v in arrays.items()})
    team_a = (torch.arange(n, device="cuda") % 2).to(torch.uint8)
    team_b = team_a.clone()
    rs_a = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    rs_b = rs_a.clone()
    reward_b = torch.zeros(n, dtype=torch.float64, device="cuda")
    done_b = torch.zeros(n, dtype=torch.uint8, device="cuda")
    turn = torch.zeros((), dtype=torch.int32, device="cuda")
    resets_a = torch.zeros(1, dtype=torch.int32, device="cuda")
    resets_b = resets_a.clone()
    for t in range(70):
        a.gen_actions(99, 0)
        b.unit_actions.copy_(a.unit_actions)
        b.tile_actions.copy_(a.tile_actions)
        a.step()
        b.step()
        if t % 9 == 4:                    # finish a few matches early so that resets happen
            for g in (a, b):
                g.hdr[t::37, 24] = 1
        reward_a = a.reward(team_a, rs_a)
        done_a = a.hdr[:, 24] != 0
        a.reset_done(pool, t, resets_a)
        rs_a.masked_fill_(done_a[:, None], 0)
        idx = torch.arange(n, dtype=torch.int64, device="cuda")
        hsh = ((idx * 0x9E3779B1 + t * 0x85EBCA6B) & 0xFFFFFFFF) * 0x2C1B3C6D & 0xFFFFFFFF
        team_a = torch.where(done_a, ((hsh >> 17) & 1).to(torch.uint8), team_a)
        turn.fill_(t)
        b.env_finish(pool, team_b, rs_b, reward_b, done_b, turn, resets_b)
        assert torch.equal(reward_a, reward_b) and torch.equal(done_a, done_b.bool()), t
        assert torch.equal(team_a, team_b) and torch.equal(rs_a, rs_b), t
        for k in a.sections():
            assert torch.equal(a.sections()[k], b.sections()[k]), (t, k)
        # offsets on the device vs on the host
        want = a.observe_packed(team_a)[3].clone()
        for cap in (100000, int(want[-1]) // 2):
            got = b.observe_packed(team_b, row_cap=cap)[3]
            assert torch.equal(got, want.clamp(max=cap)), (t, cap)
    assert int(resets_a) == int(resets_b) > 0
    assert int(b.row_overflow) > 0


def test_graphed_rollout_turns():
    """BatchedLuxEnvironment.graphed(policy): the whole turn replayed as one CUDA graph.  The observation rows it
    leaves for the next turn equal a fresh eager observation of the same state, finished matches are reset, no
    match raises a status flag, and a second graph with the same seeds reproduces the rewards."""
    from luxpythonenvgym_b200.env import BatchedLuxEnvironment

    def run(turns):
        torch.manual_seed(5)
        env = BatchedLuxEnvironment(512, size=16, seed=300, pool_maps=32, packed=True)
        env.reset()
        net = torch.nn.Sequential(torch.nn.Linear(85, 32), torch.nn.Tanh(), torch.nn.Linear(32, 9)).cuda()

        @torch.no_grad()
        def policy(obs):
            return net(obs).argmax(dim=-1).to(torch.int32)

        roll = env.graphed(policy, row_cap=8 * 512)
        total, finished = torch.zeros(512, dtype=torch.float64, device="cuda"), 0
        for t in range(turns):
            reward, done = roll.turn()
            total += reward
            finished += int(done.sum())
        return env, roll, total, finished

    env, roll, total, finished = run(80)
    g = env.game
    assert finished > 0 and int(roll.decisions) > 80 * 512 // 2 and int(g.row_overflow) == 0
    assert (g.headers()["status"] == 0).all() and (g.headers()["done"] == 0).all()
    want_obs, want_ent, want_cnt = [t.clone() for t in g.observe(env.team)]
    off = roll.offsets.cpu().numpy()
    cnt = roll.count.cpu().numpy()
    assert np.array_equal(cnt, want_cnt.cpu().numpy())
    for m in range(0, 512, 5):
        c, o = int(cnt[m]), int(off[m])
        assert torch.equal(roll.obs[o:o + c], want_obs[m, :c]) and torch.equal(roll.ent[o:o + c], want_ent[m, :c])
    _, _, total2, finished2 = run(80)
    assert finished2 == finished and torch.equal(total, total2)


def test_step_in_a_cuda_graph_equals_the_eager_steps():
    """One turn (step + fused reset / action-stream pass) captured into a CUDA graph and replayed 60 times against
    the same 60 turns issued eagerly from the same snapshot: byte-equal batches.  The classification pass's counters
    rotate through HOST state, which a replay does not advance -- a captured step has to clear its counters itself,
    or its big list and play-order lists grow from replay to replay (matches played twice).  Sparse matches with a
    dense tail so that all three launches of a step carry work; then eager steps again on the batch that was graphed."""
    n, size = 2048, 32
    z = G.npz("dense")
    caps = S.default_caps(size)
    dense = [_recap(G.get_state(z, nm + "/init"), caps) for nm in G.dense_names() if int(z[nm + "/size"]) == size]
    g = _game(n, size)
    pool = _game(64, size)
    pool.load_arrays(mapgen.generate_batch(np.arange(41000, 41064), size))
    g.load_arrays(mapgen.generate_batch(np.arange(42000, 42000 + n), size))
    g.load_states([dense[i % len(dense)] for i in range(192)], first=n - 192)
    resets = torch.zeros(1, dtype=torch.int32, device=g.device)
    g.gen_actions(77, 0)
    for t in range(30):
        g.step()
        g.reset_gen(pool, 5, 77, 0, reset_count=resets)
    torch.cuda.synchronize()
    snap = {k: v.clone() for k, v in g.sections().items()}
    ua, ta = g.unit_actions.clone(), g.tile_actions.clone()

    def rewind():
        for k, v in g.sections().items():
            v.copy_(snap[k])
        g.unit_actions.copy_(ua)
        g.tile_actions.copy_(ta)

    for t in range(60):                      # eager
        g.step()
        g.reset_gen(pool, 5, 77, 0, reset_count=resets)
    torch.cuda.synchronize()
    want = {k: v.clone() for k, v in g.sections().items()}
    rewind()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            g.step()
            g.reset_gen(pool, 5, 77, 0, reset_count=resets)
    torch.cuda.current_stream().wait_stream(side)
    rewind()                                 # (capture does not execute; start from the snapshot anyway)
    for t in range(60):
        graph.replay()
    torch.cuda.synchronize()
    for k, v in g.sections().items():
        assert torch.equal(want[k], v), "graph replays differ from the eager steps in " + k
    assert (g.headers()["status"] == 0).all()
    rewind()
    for t in range(60):                      # eager again, on a batch whose counters the replays have used
        g.step()
        g.reset_gen(pool, 5, 77, 0, reset_count=resets)
    torch.cuda.synchronize()
    for k, v in g.sections().items():
        assert torch.equal(want[k], v), "eager steps after a capture differ in " + k


def test_spill_list_and_scheduling_paths_under_load():
    """Force the rarely taken launch paths on a real batch: a tiny small-class slice (3200 B) so that a large
    share of a dense batch goes to the big class every turn, with and without overlapping the two classes,
    forced two-class mode (no automatic fallback), and more batches than the library caches scratch for."""
    lib = _lib.load()
    z = G.npz("dense")
    names = [nm for nm in G.dense_names() if int(z[nm + "/size"]) == 32]
    caps = S.default_caps(32)
    inits = [_recap(G.get_state(z, nm + "/init"), caps) for nm in names]
    sparse = mapgen.generate_batch(np.arange(4000, 4096), 32)
    try:
        # heavy 0: serialized launches; 1: round 1's dependent-launch overlap; 2: the big class forked to the side
        # stream (the default); 3: the same on a batch smaller than the machine (every CTA resident at once);
        # 4: forked, with a big-class slice so small that most dense matches go to the oversize launch (on its own
        # stream); 5: the same with the oversize launch behind the big class on its stream; 6: the small class in
        # 8-warp CTAs with a rendezvous at all nine phase boundaries (CTAs with idle warps: big-class matches, the
        # batch's ragged end); 7: the same, 5 warps, default rendezvous points
        for heavy in (0, 1, 2, 3, 4, 5, 6, 7):
            for k, v in ((b"slice", 3200), (b"two_class", 1), (b"overlap", int(heavy > 0)), (b"fork", int(heavy >= 2)),
                         (b"group", (8, 16, 32, 32, 32, 32, 32, 32)[heavy]), (b"mid_slice", 6144 if heavy in (4, 5) else 0),
                         (b"over_own", (-1, -1, -1, -1, 1, 0, -1, -1)[heavy]), (b"cta_sync", (0, 0, 0, 0, 0, 0, 8, 5)[heavy]),
                         (b"sync_mask", 0x1FF if heavy == 6 else 0x54)):
                assert lib.lux_b200_set_option(k, v) == 0
            n = 384 if heavy < 3 else 120
            g = _game(n, 32)
            g.load_states([inits[i % len(inits)] for i in range(n - 96)])       # dense: spill
            g.load_arrays(sparse, first=n - 96)                                   # fresh maps: fit
            host = gpu_util.HostBatch(g)
            for turn in range(45):
                g.gen_actions(5150 + heavy, 0)
                host.gen_actions(5150 + heavy, 0)
                g.step()
                host.step()
                bad = gpu_util.diff_batches(host, g)
                assert len(bad) == 0, "heavy=%d turn %d: %s" % (heavy, turn, gpu_util.explain(host, g, int(bad[0])))
        # more live batches than cached scratch slots (8): eviction and re-creation stay correct
        games = [_game(16, 12) for _ in range(11)]
        hosts = []
        for i, gm in enumerate(games):
            gm.load_arrays(mapgen.generate_batch(np.arange(100 * i, 100 * i + 16), 12))
            hosts.append(gpu_util.HostBatch(gm))
        for turn in range(12):
            for i, gm in enumerate(games):
                gm.gen_actions(1 + i, 0)
                hosts[i].gen_actions(1 + i, 0)
                gm.step()
                hosts[i].step()
        for i, gm in enumerate(games):
            assert len(gpu_util.diff_batches(hosts[i], gm)) == 0, i
    finally:
        for k, v in ((b"slice", 0), (b"two_class", -1), (b"overlap", 1), (b"fork", 1),
                     (b"group", 0), (b"mid_slice", 0), (b"over_own", -1), (b"cta_sync", -1), (b"sync_mask", 0x54)):
            lib.lux_b200_set_option(k, v)
