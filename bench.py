#!/usr/bin/env python
"""bench.py -- match-turns/sec of the batched Lux AI 2021 turn engine on N B200s.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W          our arm (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N ...          CPU arm: the oracle port on the host cores

A "step" is one turn of every resident match: lux_b200_step on this turn's canonical seeded
actions, then lux_b200_reset_gen (finished matches are replaced from a pool of pre-generated
maps and the next turn's actions generated, one fused pass).  `value` counts only LIVE match-turns.
Workload: BASELINE.json config 5's per-GPU shape -- 32x32 maps, 16384 resident matches per
GPU (state 230+ MB > the 126 MB L2, so no L2 flush is needed between steps), seeded random
actions.  Matches shard independently over ranks (weak scaling); the only collective is
the NCCL all-reduce of the episode-statistics vector.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZE = 32
MATCHES_PER_GPU = 16384
POOL_MAPS = 1024
BURN_IN_TURNS = 400          # untimed: brings the batch to a steady mix of match ages
STREAM_SEED = 20211019
METRIC = "match-turns/sec"
MIN_TIMED_SECONDS = 0.25     # the K-step block is repeated until this much device time is covered
JSON_OUT = sys.stdout


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--matches", type=int, default=MATCHES_PER_GPU)
    ap.add_argument("--size", type=int, default=SIZE)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = as many as --steps")
    ap.add_argument("--e2e-halves", action="store_true", help="end-to-end leg ALSO with the batch driven as two pipelined half-batches "
                                                              "(reported as e2e.two_halves_value; measured +2.4 %: the leg is device-bound)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the dense and config-4 legs (N = 1 only)")
    return ap.parse_args()


def config_dict(args, n_gpus):
    return {"workload": "BASELINE.json config 5 per-GPU shape: %dx%d maps, %d resident matches per GPU, "
                        "canonical seeded random action stream, finished matches reset from a %d-map pool"
                        % (args.size, args.size, args.matches, POOL_MAPS),
            "map_size": args.size, "matches_per_gpu": args.matches, "n_gpus": n_gpus,
            "burn_in_turns": BURN_IN_TURNS, "stream_seed": STREAM_SEED,
            "l2": "state+actions per GPU exceed the 126 MB L2 (inputs larger than L2, no flush)",
            "parallelism": "independent matches sharded over ranks (dp%d), NCCL all-reduce of stats only" % n_gpus}


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        try:  # in-process NVML: a few hundred samples even in a sub-second timed region
            import pynvml
            pynvml.nvmlInit()
            hnd = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(hnd, pynvml.NVML_CLOCK_SM)
            bits = ((pynvml.nvmlClocksThrottleReasonHwSlowdown, 2), (pynvml.nvmlClocksThrottleReasonHwThermalSlowdown, 3),
                    (pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, 4), (pynvml.nvmlClocksThrottleReasonSwPowerCap, 5))
            while not self.stop_flag:
                sm = pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM)
                reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)
                row = [str(sm), str(mx), "", "", "", ""]
                for bit, col in bits:
                    row[col] = "Active" if reasons & bit else "Not Active"
                self.rows.append(row)
                time.sleep(0.002)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.1)

    def sample_now(self):
        """One synchronous sample (called right before / after / between the timed blocks, so that even a
        millisecond-long region has clock evidence taken under load)."""
        try:
            import pynvml
            pynvml.nvmlInit()
            hnd = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            sm = pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(hnd, pynvml.NVML_CLOCK_SM)
            reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)
            row = [str(sm), str(mx), "", "", "", ""]
            for bit, col in ((pynvml.nvmlClocksThrottleReasonHwSlowdown, 2), (pynvml.nvmlClocksThrottleReasonHwThermalSlowdown, 3),
                             (pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, 4), (pynvml.nvmlClocksThrottleReasonSwPowerCap, 5)):
                row[col] = "Active" if reasons & bit else "Not Active"
            self.rows.append(row)
            return
        except Exception:
            pass
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout
            parts = [p.strip() for p in out.strip().split(",")]
            if len(parts) >= 6:
                self.rows.append(parts)
        except Exception:
            pass

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = max(int(r[1]) for r in self.rows if r[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.rows)}


def make_pool_arrays(size, count, seed0=424242):
    """`count` distinct seeded maps as host section arrays (native generator, all host threads)."""
    from luxpythonenvgym_b200 import mapgen
    return mapgen.generate_batch(np.arange(seed0, seed0 + count), size)


# ------------------------------------------------------------------------------------------------
class CpuPort:
    """The oracle port (oracle/lux_oracle.c) on `threads` host threads over a bounded sample of
    the same workload: same maps, same action stream, same pool reset rule.  Only the step
    calls are timed (action construction is not), one C call per thread per `run`."""

    def __init__(self, size, n_matches, threads, pool_arrays):
        from luxpythonenvgym_b200 import state as S
        from oracle import coracle
        self.coracle = coracle
        coracle.lib()
        self.n, self.threads = n_matches, threads
        caps = S.default_caps(size)

        def alloc(k):
            return {name: np.zeros((k, width), dtype=np.uint8) for name, width in (
                ("hdr", 128), ("cells", size * size * 8), ("units", caps["units"] * 16),
                ("cities", caps["cities"] * 16), ("tiles", caps["tiles"] * 2), ("res", caps["res"] * 4))}

        n_pool = pool_arrays["hdr"].shape[0]
        self.arr, self.parr = alloc(n_matches), {k: np.ascontiguousarray(v) for k, v in pool_arrays.items()}
        for name in self.arr:
            self.arr[name][:] = self.parr[name][np.arange(n_matches) % n_pool]
        self.desc = coracle.batch_desc(size, caps, self.arr, n_matches)
        self.pdesc = coracle.batch_desc(size, caps, self.parr, n_pool)
        self.ua = np.zeros((n_matches, caps["units"]), dtype=np.uint32)
        self.ta = np.zeros((n_matches, caps["tiles"]), dtype=np.uint8)
        self.epoch = 0

    def run(self, turns):
        """-> (match_turns_per_sec, live_match_turns, step_seconds of the slowest thread)."""
        from concurrent.futures import ThreadPoolExecutor
        chunks = [(t * self.n // self.threads, (t + 1) * self.n // self.threads) for t in range(self.threads)]

        def work(rng):
            lo, hi = rng
            return self.coracle.run_batch(self.desc, self.pdesc, STREAM_SEED, 0, self.ua, self.ta, lo, hi - lo,
                                          turns, self.epoch)

        with ThreadPoolExecutor(self.threads) as ex:
            res = list(ex.map(work, chunks))
        self.epoch += turns
        live = sum(r[1] for r in res)
        secs = max(r[0] for r in res)
        return live / secs, live, secs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_matches, turns_per_step = 256 * threads, 40
    port = CpuPort(args.size, n_matches, threads, make_pool_arrays(args.size, POOL_MAPS))
    port.run(BURN_IN_TURNS)
    vals, budget = [], 150.0
    t_begin = time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, live, secs = port.run(turns_per_step)
        if i >= args.warmup:
            vals.append((live, secs))
        if time.perf_counter() - t_begin > budget and len(vals) >= 3:
            break
    value = sum(l for l, _ in vals) / sum(s for _, s in vals)
    sample = ("oracle/lux_oracle.c: %d matches x %d turns per step (after %d burn-in turns, pool resets on), "
              "%d host threads, step calls only" % (n_matches, turns_per_step, BURN_IN_TURNS, threads))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "match-turns/s", "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([s for _, s in vals])),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config_dict(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": "match-turns/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "match-turns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference is pure Python (no compiled path) and /root/reference does not travel to the GPU "
                    "box, so this arm times the C port of the same algorithm; the unmodified Python engine measured "
                    "1-8 k match-turns/s per core in the build container (BASELINE.md section 2)"}
    print(json.dumps(line), file=JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
def dense_leg(dev, n, peak, steps=20, min_seconds=0.15):
    """Extra leg (N = 1): the same step on a DENSE late-game batch -- n resident 32x32 matches started from the
    replay states of tests/golden/dense.npz (Kaggle episodes imported at turn 80 / 200 / 300 by the reference's
    process_updates), canonical action stream, finished matches reset from the same pool.  Same timing rules as
    the main leg (blocks of `steps` steps, median block)."""
    import torch
    from luxpythonenvgym_b200 import state as S
    from luxpythonenvgym_b200.batched import BatchedGame
    from tests import golden_util as G
    z = G.npz("dense")
    caps = S.default_caps(32)
    states = []
    for name in G.dense_names():
        if int(z[name + "/size"]) != 32:
            continue
        ms = G.get_state(z, name + "/init")
        o = S.MatchState(32, caps)
        o.hdr, o.cells = ms.hdr.copy(), ms.cells.copy()
        for sec in ("units", "cities", "tiles", "res"):
            dst, src = getattr(o, sec), getattr(ms, sec)
            k = min(len(dst), len(src))
            dst[:k] = src[:k]
        states.append(o)
    game, pool = BatchedGame(n, 32, device=dev), BatchedGame(len(states), 32, device=dev)
    pool.load_states(states)
    game.load_states([states[i % len(states)] for i in range(n)])
    resets = torch.zeros(1, dtype=torch.int32, device=dev)
    seed, epoch = 31337, 0
    game.gen_actions(seed, 0)
    for _ in range(30):
        game.step()
        game.reset_gen(pool, epoch, seed, 0, reset_count=resets)
        epoch += 1
    torch.cuda.synchronize()
    blocks, spent = [], 0.0
    counters = torch.zeros((402, 2), dtype=torch.int64, device=dev)
    game.gen_actions(seed, 0, counters=counters[0])
    while len(blocks) < 400 and (spent < min_seconds or len(blocks) < 3):
        r = len(blocks)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        for k in range(steps):
            ev[k][0].record()
            game.step()
            ev[k][1].record()
            game.reset_gen(pool, epoch, seed, 0, reset_count=resets, counters=counters[r] if k < steps - 1 else counters[r + 1])
            epoch += 1
        t1.record()
        torch.cuda.synchronize()
        blocks.append((t0.elapsed_time(t1), sum(a.elapsed_time(b) for a, b in ev) / steps, r))
        spent += blocks[-1][0] * 1e-3
    blocks.sort()
    total_ms, step_ms, r = blocks[(len(blocks) - 1) // 2]
    live, alg = [int(x) for x in counters[r].cpu().tolist()]
    st = game.stats().cpu().tolist()
    achieved = alg / steps / (step_ms * 1e-3) / 1e9
    return {"workload": "%d resident 32x32 matches from the dense replay-state pool (tests/golden/dense.npz, %d states), "
                        "canonical action stream, pool resets" % (n, len(states)),
            "value": live / (total_ms * 1e-3), "unit": "match-turns/s", "ms_per_step": total_ms / steps,
            "kernel_ms": {"lux_step_kernel": step_ms, "whole_step": total_ms / steps}, "blocks": len(blocks),
            "steps_per_block": steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes_per_launch": alg / steps},
            "mean_units_per_match": st[7] / max(st[0], 1), "mean_city_tiles_per_match": st[8] / max(st[0], 1),
            "live": st[0], "status_nonzero": st[10]}


def config4_leg(dev, n_envs, steps=200):
    """Extra leg (N = 1): BASELINE.json config 4, the LuxEnvironment-shaped rollout loop on the device --
    packed per-entity observation rows -> torch MLP 85-64-64-9 (tanh; the SB3 MlpPolicy shape) sampling one
    code per actionable entity -> lux_b200_encode_actions -> lux_b200_step(prevalidate) -> lux_b200_reward ->
    masked resets and team draws.  One host synchronisation per turn (the packed row count)."""
    import torch
    from luxpythonenvgym_b200.env import BatchedLuxEnvironment
    torch.manual_seed(0)
    env = BatchedLuxEnvironment(n_envs, size=32, seed=9000, pool_maps=256, device=dev, packed=True)
    net = torch.nn.Sequential(torch.nn.Linear(85, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                              torch.nn.Linear(64, 9)).to(dev)

    @torch.no_grad()
    def act(obs):
        return torch.multinomial(torch.softmax(net(obs), dim=-1), 1).squeeze(1).to(torch.int32)

    env.reset()
    mode = "one CUDA graph per turn (policy, encode, step, reward, resets, observe); no host synchronisation"
    try:
        roll = env.graphed(act, row_cap=3 * n_envs)
        for _ in range(100):
            roll.turn()
        torch.cuda.synchronize()
        d0 = int(roll.decisions.item())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            roll.turn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        decisions = int(roll.decisions.item()) - d0
        rows = roll.row_cap * steps
        if int(env.game.row_overflow.item()):
            mode += "; %d rows cut by the row capacity" % int(env.game.row_overflow.item())
    except Exception as exc:                                    # capture refused: the eager loop, one sync per turn
        mode = "eager loop, one host synchronisation per turn (graph capture failed: %s)" % (str(exc)[:120],)
        obs, ent, cnt, offsets = env.reset()
        for _ in range(100):
            (obs, ent, cnt, offsets), reward, done = env.step(act(obs))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        decisions = torch.zeros((), dtype=torch.int64, device=dev)
        rows = 0
        e0.record()
        for _ in range(steps):
            rows += int(obs.shape[0])
            decisions += cnt.sum()
            (obs, ent, cnt, offsets), reward, done = env.step(act(obs))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        decisions = int(decisions.item())
    return {"workload": "%d envs (32x32), learning team drawn per reset, idle opponent, torch MLP policy 85-64-64-9" % n_envs,
            "ms_per_turn": ms / steps, "env_turns_per_s": n_envs * steps / (ms * 1e-3),
            "entity_decisions_per_s": decisions / (ms * 1e-3), "decisions_per_turn": decisions / steps,
            "policy_rows_per_turn": rows / steps, "steps": steps, "mode": mode,
            "obs_bytes_per_turn": 340 * decisions / steps}


def e2e_two_halves(game, pool, snap, base, epoch0, resets, W_e, E, barrier):
    """The end-to-end leg with the resident batch driven as TWO half-batches, each through its own
    submit / wait pair on its own stream: while the host blocks on one half's done flags and uploads that half's next
    action list, the device plays the other half.  Per half and per turn nothing changes: the turn's host action list
    is uploaded from pinned memory, the host waits for THAT turn's done flags + summary before it submits the half's
    next turn.  Returns (seconds for E turns of both halves, h2d bytes per turn, d2h bytes per turn); the final
    headers of the host path are checked against the device path of the same half-batches."""
    import torch
    from luxpythonenvgym_b200.batched import BatchedGame
    dev, n, size = game.device, game.n, game.size
    h = n // 2
    halves, fin, ents, ent_ks = [], [], [], []
    for i in range(2):
        g = BatchedGame(h, size, caps=game.caps, device=dev)
        for k, t in g.sections().items():
            t.copy_(snap[k][i * h:(i + 1) * h])
        # recording pass (untimed): the device action stream of this half, turn by turn, packed into entry lists
        ent_dev, ent_k = [], []
        for k in range(W_e + E):
            g.unit_actions.zero_()
            g.tile_actions.zero_()
            g.gen_actions(STREAM_SEED, base + i * h)
            ent = g.pack_entries()
            ent_dev.append(ent)
            ent_k.append(int(ent.shape[0]))
            g.step()
            g.reset_done(pool, epoch0 + k, resets)
        torch.cuda.synchronize()
        fin.append(g.hdr.clone())
        off = np.concatenate([[0], np.cumsum(ent_k)]).astype(np.int64)
        ent_h = torch.empty((max(int(off[-1]), 1), 2), dtype=torch.int32).pin_memory()
        for k, ent in enumerate(ent_dev):
            ent_h[off[k]: off[k + 1]].copy_(ent)
        del ent_dev
        for k, t in g.sections().items():     # rewind
            t.copy_(snap[k][i * h:(i + 1) * h])
        g.unit_actions.zero_()
        g.tile_actions.zero_()
        halves.append(g)
        ents.append([ent_h[off[k]: off[k + 1]] for k in range(W_e + E)])
        ent_ks.append(ent_k)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    done_h = [torch.empty(h, dtype=torch.uint8).pin_memory() for _ in range(2)]
    summ_h = [torch.zeros(4, dtype=torch.int32).pin_memory() for _ in range(2)]

    def submit(i, k):
        with torch.cuda.stream(streams[i]):
            halves[i].step_host_sparse_submit(ents[i][k], ent_ks[i][k])
            halves[i].reset_done(pool, epoch0 + k, resets)       # queued before the host blocks on this half

    def wait(i):
        with torch.cuda.stream(streams[i]):
            halves[i].step_host_wait(done_h[i], summ_h[i])

    for k in range(W_e):                     # untimed: first-call allocations, pinned-page faults
        submit(0, k); submit(1, k); wait(0); wait(1)
    barrier()
    w0 = time.perf_counter()
    submit(0, W_e); submit(1, W_e)
    for k in range(W_e + 1, W_e + E):
        wait(0); submit(0, k)
        wait(1); submit(1, k)
    wait(0); wait(1)
    torch.cuda.synchronize()
    secs = time.perf_counter() - w0
    for i in range(2):
        assert torch.equal(halves[i].hdr, fin[i]), "host-buffer path (two halves) diverged from the device path"
    h2d = sum(8 * ent_ks[i][k] for i in range(2) for k in range(W_e, W_e + E)) // max(E, 1)
    return secs, h2d, 2 * (h + 16)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from luxpythonenvgym_b200.batched import BatchedGame

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    n, size = args.matches, args.size
    # every resident match starts on its own seeded map (seed = 10^6 * rank + match index); finished
    # matches are replaced from a 1024-map pool.  Host generation: native generator, all host threads.
    t_gen = time.perf_counter()
    pool_arrays = make_pool_arrays(size, POOL_MAPS, seed0=424242 + 100000 * rank)
    game = BatchedGame(n, size, device=dev)
    pool = BatchedGame(POOL_MAPS, size, device=dev)
    pool.load_arrays(pool_arrays)
    game.load_arrays(make_pool_arrays(size, n, seed0=1000000 * (rank + 1)))
    mapgen_s = time.perf_counter() - t_gen
    base = rank * n
    resets = torch.zeros(1, dtype=torch.int32, device=dev)
    epoch = [0]

    # one step = this turn's canonical actions (generated by the fused reset + action-stream pass that closed
    # the previous step) -> lux_b200_step -> finished matches replaced and the next turn's actions generated
    def one_step(count=None):
        game.step()
        game.reset_gen(pool, epoch[0], STREAM_SEED, base, reset_count=resets, counters=count)
        epoch[0] += 1

    game.gen_actions(STREAM_SEED, base)
    for _ in range(BURN_IN_TURNS):
        one_step()
    for _ in range(max(args.warmup, 3)):
        one_step()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- timed region: blocks of EXACTLY K steps, device-timed (CUDA events on the launching stream), each
    # block bracketed by barrier + synchronize; the block is repeated R times so that at least
    # MIN_TIMED_SECONDS of device time are covered whatever K is (a 20-step block lasts 3 ms), and the MEDIAN
    # block is reported.  Per-step events around lux_b200_step give the step kernel's own duration.
    K = args.steps

    def timed_block(count_this, count_next):
        ev_a = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        ev_b = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler.sample_now()
        t0.record()
        for k in range(K):
            ev_a[k].record()
            game.step()
            ev_b[k].record()
            # the pass that closes step k counts the live matches / algorithmic bytes of the NEXT turn
            game.reset_gen(pool, epoch[0], STREAM_SEED, base, reset_count=resets,
                           counters=count_this if k < K - 1 else count_next)
            epoch[0] += 1
        t1.record()
        barrier()
        sampler.sample_now()
        return t0.elapsed_time(t1), sum(x.elapsed_time(y) for x, y in zip(ev_a, ev_b)) / K

    sampler = ClockSampler(local)
    sampler.start()
    # a first block sizes R (the same on every rank: the slowest rank's block time decides)
    R_MAX = 400
    counters = torch.zeros((R_MAX + 2, 2), dtype=torch.int64, device=dev)
    game.gen_actions(STREAM_SEED, base, counters=counters[0])      # untimed: counts the first timed turn
    launches0 = game.lib.lux_b200_launch_count()
    blocks = [timed_block(counters[0], counters[1])]
    first = torch.tensor([blocks[0][0]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(first, op=dist.ReduceOp.MAX)
    R = int(min(R_MAX, max(3, np.ceil(MIN_TIMED_SECONDS * 1e3 / max(float(first.item()), 1e-3)))))
    for r in range(1, R):
        blocks.append(timed_block(counters[r], counters[r + 1]))
    launches = (game.lib.lux_b200_launch_count() - launches0) // R
    sampler.stop_flag = True
    # the median block by whole-block time (max over ranks per block first)
    block_ms = torch.tensor([bm for bm, _ in blocks], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(block_ms, op=dist.ReduceOp.MAX)
    order = torch.argsort(block_ms)
    mid = int(order[(R - 1) // 2].item())
    total_ms = float(block_ms[mid].item())
    step_ms = blocks[mid][1]
    all_ms = sorted(block_ms.cpu().tolist())
    live_turns, alg_bytes = [int(x) for x in counters[mid].cpu().tolist()]

    # ---- end to end through the host entry point: pinned host action LIST in, done flags + summary out.
    # The device path above is replayed from a snapshot: the action lists of W_e + E turns are recorded on the
    # host first (untimed), then the same turns run through lux_b200_step_host_sparse -- W_e untimed calls
    # (first-call allocations, pinned-page faults) and E timed ones.
    E = min(args.e2e_steps, K) if args.e2e_steps > 0 else K
    W_e = 5
    snap = {k: t.clone() for k, t in game.sections().items()}
    epoch0 = epoch[0]
    ent_dev, ent_k = [], []
    for k in range(W_e + E):
        game.unit_actions.zero_()
        game.tile_actions.zero_()
        game.gen_actions(STREAM_SEED, base)
        ent = game.pack_entries()
        ent_dev.append(ent)
        ent_k.append(int(ent.shape[0]))
        game.step()
        game.reset_done(pool, epoch[0], resets)
        epoch[0] += 1
    torch.cuda.synchronize()
    final_hdr_device_path = game.hdr.clone()
    ent_off = np.concatenate([[0], np.cumsum(ent_k)]).astype(np.int64)
    ent_h = torch.empty((max(int(ent_off[-1]), 1), 2), dtype=torch.int32).pin_memory()     # ragged: one slice per turn
    for k, ent in enumerate(ent_dev):
        ent_h[ent_off[k]: ent_off[k + 1]].copy_(ent)
    del ent_dev
    for k, t in game.sections().items():     # rewind
        t.copy_(snap[k])
    epoch[0] = epoch0
    game.unit_actions.zero_()
    game.tile_actions.zero_()
    done_h = torch.empty(n, dtype=torch.uint8).pin_memory()
    summ_h = torch.zeros(4, dtype=torch.int32).pin_memory()
    h2d_total = 0
    e2e_s = 0.0
    views = [ent_h[ent_off[k]: ent_off[k + 1]] for k in range(W_e + E)]      # one host list per turn
    for k in range(W_e + E):
        if k == W_e:
            barrier()
            w0 = time.perf_counter()
        # H2D: this turn's action list (8 bytes per action); step; the pool resets of the matches that just finished
        # are queued behind it BEFORE the host blocks; D2H: done flags + summary; the host waits for them every step
        game.step_host_sparse_submit(views[k], ent_k[k])
        game.reset_done(pool, epoch[0], resets)
        game.step_host_wait(done_h, summ_h)
        if k >= W_e:
            h2d_total += 8 * ent_k[k]
        epoch[0] += 1
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    assert torch.equal(game.hdr, final_hdr_device_path), "host-buffer path diverged from the device path"
    # live match-turns of those E steps = the same trajectory the device path ran
    h2d = h2d_total // max(E, 1)
    d2h = n + 16
    # the same leg with the batch driven as two pipelined half-batches (each with its own submit / wait)
    e2e_halves_s = 0.0
    if n >= 2048 and n % 2 == 0 and args.e2e_halves:
        for k, t in game.sections().items():     # rewind (the halves start from the same snapshot)
            t.copy_(snap[k])
        pipe_s, pipe_h2d, pipe_d2h = e2e_two_halves(game, pool, snap, base, epoch0, resets, W_e, E, barrier)
        e2e_halves_s = pipe_s

    # ---- reduce over ranks: time = max, work = sum; stats vector all-reduced with NCCL
    stats = game.stats().clone()
    red = torch.tensor([total_ms, step_ms, e2e_s, e2e_halves_s], dtype=torch.float64, device=dev)
    work = torch.tensor([live_turns, alg_bytes, n * E], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    total_ms, step_ms, e2e_s, e2e_halves_s = red.cpu().tolist()
    live_all, bytes_all, e2e_turns = work.cpu().tolist()
    st = stats.cpu().tolist()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        value = live_all / (total_ms * 1e-3)
        # per-GPU: bytes of one launch on one GPU / that launch's duration
        achieved = (bytes_all / world / K) / (step_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": "match-turns/s", "n_gpus": world, "steps": K,
                "warmup": max(args.warmup, 3), "ms_per_step": total_ms / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": config_dict(args, world),
                "gpu_launches": int(launches) * world,
                "kernel_ms": {"lux_step_kernel": step_ms, "whole_step": total_ms / K},
                "timing": {"blocks": R, "steps_per_block": K, "reported": "median block (by whole-block device time, max over ranks)",
                           "block_ms_min": all_ms[0], "block_ms_median": total_ms, "block_ms_max": all_ms[-1],
                           "timed_seconds_total": sum(all_ms) * 1e-3},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "peak_source": peak_src, "kernel": "lux_step_kernel (classify + big-class launch overlapped with the small-class launch = one step)",
                             "algorithmic_bytes_per_launch": bytes_all / world / K,
                             "live_match_turns_per_launch": live_all / world / K},
                "e2e": {"value": e2e_turns / e2e_s, "unit": "match-turns/s", "h2d_bytes_per_step": h2d * world,
                        "d2h_bytes_per_step": d2h * world, "steps": E, "warmup": 5,
                        "two_halves_value": (e2e_turns / e2e_halves_s) if e2e_halves_s > 0 else None,
                        "note": "lux_b200_step_host_sparse_submit + lux_b200_step_host_wait: pinned host action LIST (8 B per action) -> device, scatter, "
                                "step, done flags + summary -> host (mapped pinned memory), the host blocks on them every step (the pool "
                                "reset of finished matches is queued before it blocks); counts resident matches per step "
                                "(finished ones are reset the same step).  The lists are recorded beforehand (the device "
                                "action stream packed into entry form): building them from host Action objects is the "
                                "caller's cost and is not in this number.  `two_halves_value` (--e2e-halves): the resident batch driven as "
                                "TWO half-batches, each through its own submit / wait pair on its own stream, so that the device plays "
                                "one half while the host turns the other around"},
                "clocks": sampler.summary(),
                "host_mapgen": {"maps": n + POOL_MAPS, "seconds": mapgen_s, "note": "untimed setup: native seeded generator"},
                "stats": {"live": st[0], "finished_pending": st[1], "sum_units": st[7], "sum_city_tiles": st[8],
                          "status_nonzero": st[10], "resets": int(resets.item())}}
        # `traffic` is NOT measured by this run (ncu cannot run inside a timed bench): it is the DRAM byte count of
        # one step from the newest committed ncu capture, carried with the capture's name and source revision so
        # that a reader can tell whether it belongs to the kernel that was timed here
        try:
            import glob
            tr = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_r*.json")))[-1]))
            line["roofline"]["traffic"] = tr.get("dram_bytes_per_launch")
            line["roofline"]["traffic_source"] = {k: tr.get(k) for k in ("capture", "git", "kernel", "note")}
            line["roofline"]["dram_gbs_from_traffic"] = tr.get("dram_bytes_per_launch", 0) / (step_ms * 1e-3) / 1e9
        except Exception:
            pass
        if world == 1 and not args.no_extra_legs:
            # free the main leg's snapshots first
            del snap, ent_h, views
            torch.cuda.empty_cache()
            line["dense"] = dense_leg(dev, n, peak)
            line["config4"] = config4_leg(dev, n)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            port = CpuPort(size, 256 * threads, threads, pool_arrays)
            port.run(BURN_IN_TURNS)
            live_sum, sec_sum, turns_sum = 0, 0.0, 0
            while sec_sum < 1.5 and turns_sum < 40000:          # ~1.5 s x threads of CPU work
                v, live, secs = port.run(600)
                live_sum, sec_sum, turns_sum = live_sum + live, sec_sum + secs, turns_sum + 600
            line["cpu_baseline"] = {"value": live_sum / sec_sum, "unit": "match-turns/s", "cores": threads, "kind": "port",
                                    "sample": "oracle/lux_oracle.c, %d matches x %d turns after %d burn-in turns, same maps / "
                                              "action stream / pool resets, %d host threads, step calls only (%.1f s of "
                                              "stepping per thread)" % (256 * threads, turns_sum, BURN_IN_TURNS, threads, sec_sum)}
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def claim_stdout():
    """Rank 0's stdout carries exactly one JSON line, printed last.  Whatever else writes to file descriptor 1
    while the bench runs (NCCL's NCCL_DEBUG=INFO / VERSION log goes to stdout) is sent to stderr instead, so
    the caller still sees it -- NCCL_DEBUG is left as the caller set it.  Returns the stream for the JSON line."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def main():
    args = parse()
    global JSON_OUT
    JSON_OUT = claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
