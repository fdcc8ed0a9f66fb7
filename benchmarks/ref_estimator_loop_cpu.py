#!/usr/bin/env python
"""CPU cost of the REFERENCE-STRUCTURED estimator loop of EKFLeeLanded.

Runs the reference's OWN classes on the CPU -- `isaacgymenvs.PVFilter.PVFilter` and `isaacgymenvs.ahrs_ekf.EKF` (one Python
object per env, stepped in Python `for idx in range(num_envs)` loops exactly like isaacgymenvs/tasks/ekf_lee_landed.py:378-391
and :417-444) and `isaacgymenvs.controllers.Controller` (batched) -- at the reference's experiment scale (512 envs,
EKFLeeExperiments.sh:4).  This is what the batched fused kernel (ozl_ekf_lee_landed_step) replaces.  It needs the reference
checkout, so it runs in the BUILD container only (never on the GPU box); the imports go through the same namespace /
`ahrs` stub as tests/golden/make_golden.py.  Prints one JSON line.

    python benchmarks/ref_estimator_loop_cpu.py [/root/reference] [--num-envs 512] [--steps 3]
"""
import argparse
import importlib.util
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ref", nargs="?", default="/root/reference")
    ap.add_argument("--num-envs", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    sys.argv = [sys.argv[0], args.ref]
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mg.install_namespace()
    mg.stub_ahrs()
    from isaacgymenvs.PVFilter import PVFilter
    from isaacgymenvs.ahrs_ekf import EKF
    from isaacgymenvs.controllers.controller import Controller
    from isaacgymenvs.controllers.control_config import control

    n, dt = args.num_envs, 0.01
    g = torch.Generator().manual_seed(0)
    root = mg.rand_states(n, g, spread=1.0)
    acc_var = torch.tensor([1.0, 1.0, 1.0])
    pvfilters = [PVFilter(acc_var, "cpu") for _ in range(n)]
    ekfs = [EKF(frequency=1 / dt) for _ in range(n)]
    Q_state = np.tile(np.array([1.0, 0.0, 0.0, 0.0]), (n, 1))
    controller = Controller(control(), "cpu")
    pos_var = torch.tensor([1.0, 1.0, 1.0]) * 0.0000001
    pos_cnt, vel_cnt = 0.0, 37.5
    t0 = time.perf_counter()
    for step in range(args.steps):
        gyro = root[:, 10:13].numpy().astype(np.float64)
        quat_wxyz = root[:, [6, 3, 4, 5]].numpy().astype(np.float64)
        acc = (torch.randn(n, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 9.8]))
        for idx in range(n):                                             # ekf_lee_landed.py:378-391
            q = Q_state[idx] / np.linalg.norm(Q_state[idx])
            Q_state[idx] = ekfs[idx].update(q, gyr=gyro[idx], ang=quat_wxyz[idx], acc=acc[idx].numpy().astype(np.float64))
        orientation = torch.tensor(Q_state, dtype=torch.float32)
        state = torch.zeros(n, 9, 1)
        for idx in range(n):                                             # ekf_lee_landed.py:417-444
            pvfilters[idx].prediction_step(acc[idx], orientation[idx], dt=dt, sim_time=step * dt, flip_Qw=False)
            if pos_cnt * dt > 1 / 20:
                pvfilters[idx].correction_step(gps_data=root[idx, 0:3], gps_var=pos_var)
                pos_cnt = 0
            else:
                pos_cnt += 1
            if vel_cnt * dt > 1 / 75:
                pvfilters[idx].correction_step(vel_data=root[idx, 7:10], vel_var=None)
                vel_cnt = 0
            else:
                vel_cnt += 1
            state[idx] = pvfilters[idx].get_states()
        est = root.clone()
        est[:, 0:3] = state[:, 0:3, 0]
        est[:, 7:10] = state[:, 3:6, 0]
        cmd = torch.zeros(n, 4)
        cmd[:, 2] = 1.0
        controller(est, cmd)                                             # ekf_lee_landed.py:493-499
    dtw = time.perf_counter() - t0
    print(json.dumps({"what": "reference-structured EKFLeeLanded estimator loop on the CPU (the reference's PVFilter / EKF objects, "
                              "one per env, Python loops; Controller batched)",
                      "num_envs": n, "steps": args.steps, "s_per_step": dtw / args.steps,
                      "us_per_env_step": dtw / args.steps / n * 1e6, "env_steps_per_sec": n * args.steps / dtw,
                      "threads": torch.get_num_threads(), "where": "build container CPU (not the GPU box)"}))


if __name__ == "__main__":
    main()
