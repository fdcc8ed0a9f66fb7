"""Profiling aid: steady-state launches of the x500 step kernel at one size (ncu -k regex:quad_step -s <skip> -c 1 ...).
usage: python profiles/prof_step.py <n_envs> [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ouzelum_b200 import _lib  # noqa: E402
from ouzelum_b200.sim import QuadSim  # noqa: E402

n = int(sys.argv[1])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 120
dev = torch.device("cuda:0")
sim = QuadSim(_lib.default_cfg(n, fault_mode=1, seed=0), dev)
obs, rew = torch.zeros(n, 13, device=dev), torch.zeros(n, device=dev)
rs, pg = torch.ones(n, dtype=torch.int64, device=dev), torch.zeros(n, dtype=torch.int64, device=dev)
to, er = torch.zeros(n, dtype=torch.uint8, device=dev), torch.zeros(n, device=dev)
acts = [torch.rand(n, 4, device=dev) * 2 - 1 for _ in range(4)]
for k in range(steps):
    sim.step(acts[k & 3], obs, rew, rs, pg, to, er)
torch.cuda.synchronize()
print("done", n, steps)
