#!/usr/bin/env python
"""The reference's EKFLeeLanded robustness sweep (isaacgymenvs/EKFLeeExperiments.sh:4-19) on the B200 path.

The shell script runs `train.py task=EKFLeeLanded num_envs=512 test=True max_iterations=1000` once per sensor-fault setting
(flicker p in {0, 0.3, 0.4, 0.5}; random_noise sigma in {0.15, 0.20, 0.25}; flickering_and_random_noise sigma in the same
three) and each run leaves `metrics/<pomdp>_<prob>.txt` (landings) and `metrics/<pomdp>_<prob>_ep_count.txt` (resets) behind
(ekf_lee_landed.py:319-331).  In test mode the rl_games player only supplies actions the task ignores (the Lee controller
flies the vehicle), so the sweep is `max_iterations` env steps per setting.  Here every setting is one env of 512 (default)
envs stepped from a CUDA graph; the files are written once at the end from the device-side episode statistics.

    python benchmarks/ekf_lee_experiments.py [--num-envs 512] [--iterations 1000] [--out runs/ekf_lee] [--per-env-triggers]

Prints one JSON line per setting.  Reference figure to compare with: 23 landings in 26 episodes (`Landed`, flicker 0.01,
1 env; isaacgymenvs/metrics/flicker_0.01.txt) -- the reference records no EKFLeeLanded numbers.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ouzelum_b200  # noqa: E402

SWEEP = [("flicker", 0.0), ("flicker", 0.3), ("flicker", 0.4), ("flicker", 0.5),
         ("random_noise", 0.15), ("random_noise", 0.20), ("random_noise", 0.25),
         ("flickering_and_random_noise", 0.15), ("flickering_and_random_noise", 0.20), ("flickering_and_random_noise", 0.25)]


def run_setting(pomdp, prob, num_envs, iterations, out, per_env_triggers, seed=0, device="cuda:0", **over):
    cfg = ouzelum_b200.task_config("EKFLeeLanded", num_envs, seed=seed, POMDP=pomdp, pomdp_prob=prob, useCudaGraph=True,
                                   perEnvSensorTriggers=per_env_triggers, **over)
    env = ouzelum_b200.make(seed=seed, task="EKFLeeLanded", num_envs=num_envs, sim_device=device, rl_device=device,
                            headless=True, cfg=cfg)
    a = env.zero_actions()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iterations):
        env.step(a)
    landings, resets = env.write_metrics(out)
    dt = time.perf_counter() - t0
    m = env.sim.metrics().cpu()
    episodes = int(m[9])
    line = {"task": "EKFLeeLanded", "POMDP": pomdp, "pomdp_prob": prob, "num_envs": num_envs, "iterations": iterations,
            "landings": landings, "resets": resets, "episodes_finished": episodes,
            "landing_rate": (landings / episodes) if episodes else None,
            "timeouts": int(m[11]), "crash_dist": int(m[12]), "crash_z": int(m[13]),
            "sensor_triggers": "per-env" if per_env_triggers else "shared (reference)", "wall_s": dt,
            "env_steps_per_sec": num_envs * iterations / dt}
    env.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", type=int, default=512)
    ap.add_argument("--iterations", type=int, default=1000)
    ap.add_argument("--out", default="runs/ekf_lee")
    ap.add_argument("--per-env-triggers", action="store_true")
    args = ap.parse_args()
    for pomdp, prob in SWEEP:
        print(json.dumps(run_setting(pomdp, prob, args.num_envs, args.iterations, args.out, args.per_env_triggers)), flush=True)


if __name__ == "__main__":
    main()
