"""N>1 host logic on CPU: world_size-2 `gloo` process group (127.0.0.1) exercising the env sharding and the metrics
all-reduce, plus shard invariance of the counter-RNG step oracle (what makes weak/strong scaling results identical)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, steps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from ouzelum_b200 import dist as od
    from oracle.quad_step import QuadStepOracle, default_cfg
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, w, _ = od.rank_info()
    assert (r, w) == (rank, world)
    n, base = od.shard(total, rank, world)
    ora = QuadStepOracle(default_cfg(n, env_id_base=base, seed=17, fault_mode=1, max_episode_length=30))
    g = torch.Generator().manual_seed(5)
    for t in range(steps):
        a_full = torch.rand(total, 4, generator=g) * 2 - 1            # every rank draws the same global action tensor
        ora.step(a_full[base:base + n])
    m = torch.zeros(16, dtype=torch.float64)
    m[0:2] = torch.from_numpy(ora.msum[0:2])
    m[8:16] = torch.from_numpy(ora.mcnt.astype(np.float64))
    od.allreduce_metrics(m)
    torch.save({"m": m, "obs": ora.obs_buf, "reset": ora.reset_buf, "n": n, "base": base}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_covers_every_env_once():
    from ouzelum_b200.dist import shard
    for total in (1, 7, 16384, 1048576, 1000003):
        for world in (1, 2, 3, 4, 8):
            parts = [shard(total, r, world) for r in range(world)]
            assert sum(n for n, _ in parts) == total
            nxt = 0
            for n, base in parts:
                assert base == nxt
                nxt += n
    with pytest.raises(ValueError):
        shard(10, 2, 2)


def test_two_rank_gloo_sharded_oracle_matches_single_rank(tmp_path):
    from oracle.quad_step import QuadStepOracle, default_cfg
    total, steps, world = 203, 45, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, steps, str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    ref = QuadStepOracle(default_cfg(total, seed=17, fault_mode=1, max_episode_length=30))
    g = torch.Generator().manual_seed(5)
    for t in range(steps):
        ref.step(torch.rand(total, 4, generator=g) * 2 - 1)
    assert parts[0]["n"] + parts[1]["n"] == total and parts[1]["base"] == parts[0]["n"]
    assert torch.equal(torch.cat([p["obs"] for p in parts]), ref.obs_buf)          # sharding does not change any env
    assert torch.equal(torch.cat([p["reset"] for p in parts]), ref.reset_buf)
    for p in parts:                                                               # all-reduced metrics == single-rank metrics
        np.testing.assert_array_equal(p["m"][8:16].numpy(), ref.mcnt.astype(np.float64))
        np.testing.assert_allclose(p["m"][0:2].numpy(), ref.msum[0:2], rtol=1e-12)
    assert ref.mcnt[1] > 0


def test_summarize_metrics():
    from ouzelum_b200.dist import summarize
    m = torch.zeros(16, dtype=torch.float64)
    m[0], m[1], m[8], m[9], m[10] = 50.0, 30.0, 100.0, 3.0, 60.0
    s = summarize(m)
    assert s["mean_reward"] == 0.5 and s["mean_episode_return"] == 10.0 and s["mean_episode_length"] == 20.0
