"""torchrun check of the NVLink peer-memory metrics exchange (ouzelum_b200.dist.PeerMetrics, csrc/peer_metrics.cu) against an
NCCL all-reduce of the same vectors, and its cost next to it (CUDA events, max over ranks).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 benchmarks/peer_metrics_check.py
Prints one JSON line on rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ouzelum_b200 import _lib  # noqa: E402
from ouzelum_b200.dist import PeerMetrics, rank_info  # noqa: E402
from ouzelum_b200.sim import QuadSim  # noqa: E402

rank, world, local = rank_info()
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
n = 16384
sim = QuadSim(_lib.default_cfg(n, seed=rank, env_id_base=rank * n, fault_mode=1, max_episode_length=25), dev)
obs, rew = torch.zeros(n, 13, device=dev), torch.zeros(n, device=dev)
reset, prog = torch.ones(n, dtype=torch.int64, device=dev), torch.zeros(n, dtype=torch.int64, device=dev)
tout, epr = torch.zeros(n, dtype=torch.uint8, device=dev), torch.zeros(n, device=dev)
act = torch.rand(n, 4, device=dev, generator=torch.Generator(device=dev).manual_seed(rank)) * 2 - 1
x = PeerMetrics(dev)
ok = True
for rnd in range(20):
    for _ in range(7):
        sim.step(act, obs, rew, reset, prog, tout, epr)
    loc = torch.zeros(16, dtype=torch.float64, device=dev)
    x.push(sim, local=loc)
    got = x.sum(sim)
    gathered = [torch.zeros_like(loc) for _ in range(world)]
    dist.all_gather(gathered, loc)
    want = torch.zeros_like(loc)
    for g_ in gathered:                      # rank order, like the kernel
        want += g_
    ok = ok and bool(torch.equal(got, want))
    ref = loc.clone()
    dist.all_reduce(ref)
    ok = ok and bool(torch.allclose(got, ref, rtol=1e-14, atol=0))
st = x.status()
ok = ok and st["error"] == 0 and st["pushed"] == st["summed"] == 20


def timed(fn, reps=200):
    for _ in range(20):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


buf = torch.zeros(16, dtype=torch.float64, device=dev)


def peer():
    x.push(sim)
    x.sum(sim, out=buf)


def nccl():
    sim.metrics(out=buf)
    dist.all_reduce(buf)


us_peer, us_nccl = timed(peer), timed(nccl)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "peer_exchange_equals_rank_ordered_sum_and_nccl": bool(flag.item()),
                      "us_per_exchange_peer_push_plus_sum": us_peer, "us_per_exchange_metrics_read_plus_nccl_allreduce": us_nccl,
                      "note": "eager launches back to back on one stream; host launch cost included in both"}), flush=True)
dist.destroy_process_group()
