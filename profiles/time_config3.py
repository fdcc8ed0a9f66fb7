"""Timing aid: BASELINE config 3 (EKFLeeLanded, 65536 envs, DR + noise 0.15 + EKF + PV + Lee, one launch per step) replayed from a
CUDA graph.  Prints one JSON line.  Kernel variants: OUZELUM_B200_LIB=<lib>, OZL_EKF_BLOCK=128|256|512.
usage: python profiles/time_config3.py [n_envs] [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ouzelum_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
DEV = "cuda:0"
cfg = ouzelum_b200.task_config("EKFLeeLanded", n, seed=0, POMDP="random_noise", pomdp_prob=0.15, ConvergenceTime=20,
                               domainRandomization={"enable": True}, rotorFault={"enable": True},
                               exposeEstimates=bool(int(os.environ.get("OZL_EXPOSE_EST", "0"))))
env = ouzelum_b200.make(seed=0, task="EKFLeeLanded", num_envs=n, sim_device=DEV, rl_device=DEV, headless=True, cfg=cfg)
a = torch.zeros(n, 4, device=DEV)
for _ in range(60):
    env.step(a)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(steps):
        env._launch(a)
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
print(json.dumps({"lib": os.path.basename(os.environ.get("OUZELUM_B200_LIB", "in-tree")), "expose_est": os.environ.get("OZL_EXPOSE_EST", "0"), "ekf_block": os.environ.get("OZL_EKF_BLOCK", "default"),
                  "n_envs": n, "us_per_step": best, "env_steps_per_sec": n / best * 1e6, "frac_of_hbm_1372B": 1372 * n / best / 1e3 / 6552.3,
                  "episodes": env.episodes, "landings": env.landings}), flush=True)
