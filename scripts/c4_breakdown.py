import sys, time
sys.path.insert(0, "/root/repo")
import torch
from luxpythonenvgym_b200.env import BatchedLuxEnvironment
sys.path.insert(0, "/root/repo/scripts")
from bench_configs import Policy
torch.manual_seed(0)
env = BatchedLuxEnvironment(16384, size=32, seed=9000, pool_maps=256, packed=True)
policy = Policy().cuda()
obs, ent, cnt, off = env.reset()
for w in range(60):
    codes, _ = policy.act_packed(obs, ent)
    (obs, ent, cnt, off), reward, done = env.step(codes)
g = env.game
def timeit(fn, n=50):
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n, (time.perf_counter()-t0)/n*1e3
print("observe_packed (gpu ms, wall ms)", timeit(lambda: g.observe_packed(env.team)))
sel = env.team
print("observe padded", timeit(lambda: g.observe(sel, env.e_cap)))
print("policy", timeit(lambda: policy.act_packed(obs, ent)))
codes, _ = policy.act_packed(obs, ent)
print("encode", timeit(lambda: g.encode_actions(ent, cnt, codes, offsets=off)))
print("reward", timeit(lambda: g.reward(env.team, env.reward_state)))
print("draw_teams", timeit(lambda: env._draw_teams()))
print("done.any", timeit(lambda: bool((g.hdr[:, 24] != 0).any())))
print("rows", obs.shape, int(cnt.sum()))
