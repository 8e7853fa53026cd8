"""Extra measurements for BASELINE.json configs 3 and 4 (not the contract bench; see bench.py for that).

  config 3: 65536 matches of mixed map sizes 12/16/24/32 (16384 each), dense late-game populations on
            16/24/32 (fixture replay states imported at turn 80/200/300, tests/golden/dense.npz) and
            seeded-random play on 12; canonical action stream, finished matches reset from the same pool.
  config 4: LuxEnvironment-shaped PPO rollout: 16384 envs (32x32), per-entity observation tensors ->
            torch MLP policy 85-64-64-9 (tanh, the SB3 MlpPolicy shape) sampling one code per actionable
            entity -> encode -> step(prevalidate) -> reward; opponent idle.

Writes one JSON object to stdout.  Device-timed with CUDA events after warm-up.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from luxpythonenvgym_b200 import mapgen  # noqa: E402
from luxpythonenvgym_b200 import state as S  # noqa: E402
from luxpythonenvgym_b200.batched import BatchedGame  # noqa: E402
from luxpythonenvgym_b200.env import BatchedLuxEnvironment  # noqa: E402


for _opt in os.environ.get("LUX_OPTS", "").split(","):        # e.g. LUX_OPTS=fork=0,big_share=12 (tuning probes)
    if "=" in _opt:
        from luxpythonenvgym_b200 import _lib as _l
        _k, _v = _opt.split("=")
        assert _l.load().lux_b200_set_option(_k.encode(), int(_v)) == 0, _opt


def recap(ms, caps):
    o = S.MatchState(ms.size, caps)
    o.hdr, o.cells = ms.hdr.copy(), ms.cells.copy()
    for name in ("units", "cities", "tiles", "res"):
        dst, src = getattr(o, name), getattr(ms, name)
        k = min(len(dst), len(src))
        dst[:k] = src[:k]
    return o


def config3(n_per_size=16384, steps=100):
    z = np.load(os.path.join(ROOT, "tests", "golden", "dense.npz"))
    names = [str(x) for x in z["names"]]
    games = []
    for size in (12, 16, 24, 32):
        caps = S.default_caps(size)
        if size == 12:
            states = [mapgen.generate(8000 + i, size=12) for i in range(64)]
        else:
            states = []
            for nm in names:
                if int(z[nm + "/size"]) == size:
                    ms = S.MatchState.from_arrays(size, {k: z["%s/init/%s" % (nm, k)] for k in S.MatchState.SECTIONS}, legacy_res=True)
                    states.append(recap(ms, caps))
        g = BatchedGame(n_per_size, size)
        pool = BatchedGame(len(states), size)
        pool.load_states(states)
        g.load_states([states[i % len(states)] for i in range(n_per_size)])
        games.append((size, g, pool))
    counters = torch.zeros(2, dtype=torch.int64, device="cuda")

    # one stream per batch: the four batches are independent, so the tail of one size's step overlaps with the
    # next size's (LUX_C3_STREAMS=0: everything on one stream, as in round 1)
    multi = os.environ.get("LUX_C3_STREAMS", "1") != "0"
    streams = [torch.cuda.Stream() for _ in games]
    counts = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in games]

    def step_all(epoch, count):
        for i, (size, g, pool) in enumerate(games):
            with torch.cuda.stream(streams[i] if multi else torch.cuda.current_stream()):
                g.gen_actions(31337, 0, counters=counts[i] if count is not None else None)
                g.step()
                g.reset_done(pool, epoch)

    for w in range(30):
        step_all(w, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record()
    for st in streams:
        st.wait_stream(main)            # the timed region opens on every stream
    for k in range(steps):
        step_all(100 + k, counters)
    for st in streams:
        main.wait_stream(st)            # ... and closes when all of them are done
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    live, alg = [int(x) for x in sum(c.cpu() for c in counts).tolist()]
    pops = {size: {"units_per_match": float(g.stats().cpu()[7]) / g.n, "tiles_per_match": float(g.stats().cpu()[8]) / g.n,
                   "status_nonzero": int(g.stats().cpu()[10])} for size, g, pool in games}
    return {"matches": 4 * n_per_size, "steps": steps, "ms_per_step_all_sizes": ms / steps,
            "match_turns_per_s": live / (ms * 1e-3), "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "populations": pops}


class Policy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Linear(85, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                                       torch.nn.Linear(64, 9))

    @torch.no_grad()
    def act_packed(self, obs, ent):
        """Packed rows: one forward pass over all rows; unused rows (ent == -1) get code 0 and are ignored."""
        logits = self.net(obs)
        codes = torch.multinomial(torch.softmax(logits, dim=-1), 1).squeeze(1).to(torch.int32)
        return codes, int(obs.shape[0])

    @torch.no_grad()
    def act(self, obs, cnt):
        n, e_cap, _ = obs.shape
        mask = torch.arange(e_cap, device=obs.device)[None, :] < cnt[:, None]
        rows = obs[mask]
        codes = torch.zeros((n, e_cap), dtype=torch.int32, device=obs.device)
        if rows.shape[0]:
            logits = self.net(rows)
            codes[mask] = torch.multinomial(torch.softmax(logits, dim=-1), 1).squeeze(1).to(torch.int32)
        return codes, int(rows.shape[0])


def config4(n_envs=16384, steps=300):
    torch.manual_seed(0)
    env = BatchedLuxEnvironment(n_envs, size=32, seed=9000, pool_maps=256, packed=True)
    policy = Policy().cuda()
    obs, ent, cnt, offsets = env.reset()
    for w in range(150):          # the host side of this loop (torch ops, allocator) needs a longer warm-up than the kernels
        codes, _ = policy.act_packed(obs, ent)
        (obs, ent, cnt, offsets), reward, done = env.step(codes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rows, decisions = 0, torch.zeros((), dtype=torch.int64, device="cuda")
    e0.record()
    for k in range(steps):
        codes, nrows = policy.act_packed(obs, ent)
        rows += nrows
        decisions += cnt.sum()
        (obs, ent, cnt, offsets), reward, done = env.step(codes)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    decisions = int(decisions.item())
    return {"envs": n_envs, "steps": steps, "ms_per_turn": ms / steps, "env_turns_per_s": n_envs * steps / (ms * 1e-3),
            "entity_decisions_per_s": decisions / (ms * 1e-3), "decisions_per_turn": decisions / steps,
            "policy_rows_per_turn": rows / steps,
            "note": "packed observation rows; includes observation kernel, torch MLP + multinomial sampling, encode, "
                    "step(prevalidate), reward, resets"}


if __name__ == "__main__":
    out = {"config3_mixed_dense": config3(), "config4_ppo_rollout": config4()}
    print(json.dumps(out))
