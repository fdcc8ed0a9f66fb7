"""Profiling / timing aid: steady-state steps of one task class (ncu -k regex:quad_step -s <skip> -c 1 ...).
usage: python profiles/prof_task.py <Task> <n_envs> [steps]   -- prints us per step from a CUDA-graph replay"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ouzelum_b200  # noqa: E402

task, n = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
DEV = "cuda:0"
env = ouzelum_b200.make(seed=0, task=task, num_envs=n, sim_device=DEV, rl_device=DEV, headless=True,
                        cfg=ouzelum_b200.task_config(task, n, seed=0))
a = torch.rand(n, env.num_actions, device=DEV) * 2 - 1
for _ in range(60):
    env.step(a)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(steps):
        env._fused_step(a)
g.replay()
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(400_000)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) * 1e3 / steps)
print(json.dumps({"task": task, "n_envs": n, "us_per_step": best, "env_steps_per_sec": n / best * 1e6}), flush=True)
