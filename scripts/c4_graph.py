"""Config-4 turn as one CUDA graph (profiling target):  python scripts/c4_graph.py [turns] [row_cap_factor]"""
import sys

sys.path.insert(0, ".")
import torch

from luxpythonenvgym_b200.env import BatchedLuxEnvironment

turns = int(sys.argv[1]) if len(sys.argv) > 1 else 5
factor = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.manual_seed(0)
n = 16384
env = BatchedLuxEnvironment(n, size=32, seed=9000, pool_maps=256, packed=True)
net = torch.nn.Sequential(torch.nn.Linear(85, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                          torch.nn.Linear(64, 9)).cuda()


@torch.no_grad()
def act(obs):
    return torch.multinomial(torch.softmax(net(obs), dim=-1), 1).squeeze(1).to(torch.int32)


env.reset()
roll = env.graphed(act, row_cap=factor * n)
for _ in range(100):
    roll.turn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(turns):
    roll.turn()
e1.record()
torch.cuda.synchronize()
print("ms per turn %.4f  decisions %d overflow %d" % (e0.elapsed_time(e1) / turns, int(roll.decisions), int(env.game.row_overflow)))
