#!/usr/bin/env python
"""Side benchmarks for the BASELINE.json configs that are not the bench.py headline (configs[1]).
Prints one JSON line per config; run on a B200:  python benchmarks/configs.py [--only 1,3,4,5]

  config 1  Quadcopter hover, 256 envs, random actions U(-1,1) [256,12] from Generator(0), 1000 steps  (+ CPU oracle beside it)
  config 3  x500 tracking + DR + sensor noise + per-env EKF / PV filter / Lee controller, 65536 envs
  config 4  x500 rotor-fault task at the per-GPU shard sizes of the 1 Mi-env job (131072 .. 1048576 envs on this GPU)
  config 5  RPO-LSTM rollout collection (env kernels + torch LSTM policy), 32768 envs, 16-step rollouts
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# Under torchrun (config 5 on N GPUs: `python -m torch.distributed.run --nproc-per-node N benchmarks/configs.py --only 5`) every
# rank pins ITS GPU as the only visible device, so it is "cuda:0" everywhere -- the convention the reference's trainers need
# (they hard-code "cuda:0").  Must happen before torch initialises CUDA.
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
if WORLD > 1 and __name__ == "__main__":       # (bench.py imports this module for its `side_configs` and addresses GPUs by local rank)
    os.environ["CUDA_VISIBLE_DEVICES"] = os.environ.get("LOCAL_RANK", "0")
import torch  # noqa: E402

import ouzelum_b200  # noqa: E402

DEV = "cuda:0"


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def config1(DEV=DEV):
    n, steps = 256, 1000
    env = ouzelum_b200.make(seed=0, task="Quadcopter", num_envs=n, sim_device=DEV, rl_device=DEV, headless=True)
    g = torch.Generator().manual_seed(0)
    acts = [(torch.rand(n, 12, generator=g) * 2 - 1).to(DEV) for _ in range(16)]
    k = [0]

    def f():
        env.step(acts[k[0] & 15])
        k[0] += 1
    dt = timed(f, steps, 50)
    env_g = ouzelum_b200.make(seed=0, task="Quadcopter", num_envs=n, sim_device=DEV, rl_device=DEV, headless=True,
                              cfg=ouzelum_b200.task_config("Quadcopter", n, seed=0, useCudaGraph=True))

    def fg():
        env_g.step(acts[k[0] & 15])
        k[0] += 1
    dt_g = timed(fg, steps, 50)
    # the kernel alone: 1000 steps captured in one graph (what a GPU-resident consumer can reach)
    env_k = ouzelum_b200.make(seed=0, task="Quadcopter", num_envs=n, sim_device=DEV, rl_device=DEV, headless=True)
    env_k._launch(acts[0])
    torch.cuda.synchronize()
    gk = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gk):
        for j in range(steps):
            env_k._launch(acts[j & 15])
    dt_k = timed(gk.replay, 3, 1) / 3
    from oracle.quadcopter import QuadcopterOracle
    torch.set_num_threads(os.cpu_count() or 1)
    ora = QuadcopterOracle(n, seed=0)
    ca = [a.cpu() for a in acts]
    for i in range(5):
        ora.step(ca[i])
    t0 = time.perf_counter()
    cs = 100
    for i in range(cs):
        ora.step(ca[i & 15])
    cdt = time.perf_counter() - t0
    return {"config": 1, "workload": "Quadcopter hover, 256 envs, random actions, 1000 steps (per-step Python launches, no graph)",
            "env_steps_per_sec": n * steps / dt, "us_per_step": dt / steps * 1e6,
            "env_steps_per_sec_cuda_graph_per_step": n * steps / dt_g, "us_per_step_cuda_graph_per_step": dt_g / steps * 1e6,
            "env_steps_per_sec_1000_steps_in_one_graph": n * steps / dt_k, "us_per_step_1000_steps_in_one_graph": dt_k / steps * 1e6,
            "cpu_port_env_steps_per_sec": n * cs / cdt, "cpu_cores": os.cpu_count(),
            "cpu_note": "torch-CPU eager oracle of the same step (Isaac Gym CPU pipeline unavailable)"}


def config3(DEV=DEV, steps=200):
    n = 65536
    kw = dict(seed=0, POMDP="random_noise", pomdp_prob=0.15, ConvergenceTime=20, domainRandomization={"enable": True},
              rotorFault={"enable": True})
    a = torch.zeros(n, 4, device=DEV)

    def graph_us(env):
        """`steps` control steps (one launch each) replayed from ONE CUDA graph: what a GPU-resident consumer sees -- no per-step
        host work in the timed region."""
        for _ in range(40):
            env.step(a)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(steps):
                env._launch(a)
        g.replay()
        return timed(g.replay, 3, 1) / 3 / steps * 1e6
    env = ouzelum_b200.make(seed=0, task="EKFLeeLanded", num_envs=n, sim_device=DEV, rl_device=DEV, headless=True,
                            cfg=ouzelum_b200.task_config("EKFLeeLanded", n, **kw))
    us = graph_us(env)
    landings, episodes = env.landings, env.episodes
    del env
    env2 = ouzelum_b200.make(seed=0, task="EKFLeeLanded", num_envs=n, sim_device=DEV, rl_device=DEV, headless=True,
                             cfg=ouzelum_b200.task_config("EKFLeeLanded", n, perEnvSensorTriggers=True, **kw))
    us2 = graph_us(env2)
    del env2
    # the public per-step API (VecTask.step from Python, one CUDA-graph replay per call): host-launch bound at this step length
    env3 = ouzelum_b200.make(seed=0, task="EKFLeeLanded", num_envs=n, sim_device=DEV, rl_device=DEV, headless=True,
                             cfg=ouzelum_b200.task_config("EKFLeeLanded", n, useCudaGraph=True, **kw))
    dt3 = timed(lambda: env3.step(a), steps, 40)
    alg = 1372
    return {"config": 3, "workload": "x500 + DR + sensor noise sigma 0.15 + EKF (f64) + PV filter (full 9x9) + Lee controller, 65536 envs",
            "env_steps_per_sec": n / us * 1e6, "us_per_step": us, "alg_bytes_per_env_step": alg,
            "achieved_GBps_alg": alg * n / us / 1e3, "frac_of_hbm_peak": alg * n / us / 1e3 / 6552.3,
            "launches_per_step": 1, "timing": f"{steps} steps replayed from one CUDA graph, CUDA events",
            "per_env_sensor_triggers_env_steps_per_sec": n / us2 * 1e6,
            "python_step_call_env_steps_per_sec": n * steps / dt3, "python_step_call_us": dt3 / steps * 1e6,
            "landings": landings, "episodes": episodes}


def config4():
    from ouzelum_b200 import _lib
    from ouzelum_b200.sim import QuadSim
    out = []
    for n in (131072, 262144, 524288, 1048576, 4194304):
        sim = QuadSim(_lib.default_cfg(n, fault_mode=1), DEV)
        obs, rew = torch.zeros(n, 13, device=DEV), torch.zeros(n, device=DEV)
        rs, pg = torch.ones(n, dtype=torch.int64, device=DEV), torch.zeros(n, dtype=torch.int64, device=DEV)
        to, er = torch.zeros(n, dtype=torch.uint8, device=DEV), torch.zeros(n, device=DEV)
        acts = [torch.rand(n, 4, device=DEV) * 2 - 1 for _ in range(2)]
        k = [0]

        def f():
            sim.step(acts[k[0] & 1], obs, rew, rs, pg, to, er)
            k[0] += 1
        dt = timed(f, 100, 20)
        out.append({"n_envs": n, "us_per_step": dt / 100 * 1e6, "env_steps_per_sec": n * 100 / dt,
                    "alg_GBps": 284 * n * 100 / dt / 1e9, "frac_of_6552": 284 * n * 100 / dt / 1e9 / 6552.3})
        del sim
    return {"config": 4, "workload": "x500 rotor-fault step at the per-GPU shard sizes of the 1 Mi-env job (and 4 Mi)", "sizes": out}


def config5(DEV=DEV, iters=20):
    from ouzelum_b200.pomdp import POMDPWrapper
    from ouzelum_b200.rollout import RecurrentActor, RolloutStorage, collect_rollout, initial_rollout_state
    n, T = 32768, 16
    cfg = ouzelum_b200.task_config("Landing", n, seed=0, rotorFault={"enable": True}, envIdBase=RANK * n)
    env = ouzelum_b200.make(seed=0, task="Landing", num_envs=n, sim_device=DEV, rl_device=DEV, headless=True, cfg=cfg)
    actor = RecurrentActor().to(DEV)
    store = RolloutStorage(T, n, 13, 4, DEV)
    pomdp = POMDPWrapper("flicker", 0.1)
    st = [initial_rollout_state(env, actor)]

    def f():
        st[0] = collect_rollout(env, actor, store, st[0], pomdp)
    dt_fp32 = timed(f, iters, 3)
    torch.backends.cuda.matmul.allow_tf32 = True          # the policy GEMMs on tensor cores (TF32); the env kernels are unaffected
    dt = timed(f, iters, 3)
    from ouzelum_b200.rollout import GraphedRollout
    gro = GraphedRollout(env, actor, store, pomdp)
    dt_graph = timed(gro.run, iters, 3)
    # env-kernel share: the same number of env steps without the policy
    a = torch.zeros(n, 4, device=DEV)
    dte = timed(lambda: env.step(a), iters * T, 10)
    if WORLD > 1:
        # envs are sharded over the GPUs with no communication; the job's rate is set by the slowest rank (max time over ranks)
        import torch.distributed as dist
        t = torch.tensor([dt_fp32, dt, dt_graph, dte], dtype=torch.float64, device=DEV)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_fp32, dt, dt_graph, dte = (float(x) for x in t.tolist())
        n = n * WORLD
    return {"config": 5, "workload": "RPO-LSTM rollout collection: Landing task + flicker 0.1 + MLP(13-512-256)+LSTM(256-128) policy, 32768 envs per GPU, 16-step rollouts",
            "n_gpus": WORLD, "envs_total": n,
            "env_steps_per_sec": n * T * iters / dt, "us_per_env_step_call": dt / (iters * T) * 1e6,
            "env_steps_per_sec_fp32_simt_policy": n * T * iters / dt_fp32,
            "env_steps_per_sec_cuda_graph": n * T * iters / dt_graph, "us_per_env_step_call_cuda_graph": dt_graph / (iters * T) * 1e6, "policy_matmul": "TF32 tensor cores (torch.backends.cuda.matmul.allow_tf32)",
            "env_only_us_per_step": dte / (iters * T) * 1e6, "env_share_of_wall": dte / dt}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="1,3,4,5")
    args = ap.parse_args()
    fns = {"1": config1, "3": config3, "4": config4, "5": config5}
    if WORLD > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(DEV))
        if args.only != "5":
            sys.exit("multi-GPU runs of this script are for config 5 only (--only 5)")
    for k in args.only.split(","):
        out = fns[k]()
        if RANK == 0:
            print(json.dumps(out), flush=True)
    if WORLD > 1:
        dist.destroy_process_group()
