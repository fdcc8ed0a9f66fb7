"""Timing aid: CUDA-graph replay of the x500 step at several sizes, warm (one buffer set) and cold (rotating buffer sets whose
combined footprint exceeds the 126 MB L2).  Prints one JSON line per size.  usage: python profiles/time_sizes.py [sizes...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ouzelum_b200 import _lib  # noqa: E402
from ouzelum_b200.sim import QuadSim  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6552.3
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


def shard(n, seed):
    sim = QuadSim(_lib.default_cfg(n, fault_mode=1, seed=seed), dev)
    return (sim, torch.zeros(n, 13, device=dev), torch.zeros(n, device=dev), torch.ones(n, dtype=torch.int64, device=dev),
            torch.zeros(n, dtype=torch.int64, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev), torch.zeros(n, device=dev))


def graph_time(fn, count, reps=3):
    for k in range(count):
        fn(k)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k in range(count):
            fn(k)
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400_000)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / count)
    return best


sizes = [int(x) for x in sys.argv[1:]] or [16384, 65536, 131072, 262144, 524288, 1048576]
for n in sizes:
    per = n * 297
    S = max(2, int(400e6 // per) + 1)              # rotating sets: > 400 MB in total
    sh = [shard(n, j) for j in range(S)]
    a = [torch.rand(n, 4, device=dev) * 2 - 1 for _ in range(2)]

    def warm(k):
        s = sh[0]
        s[0].step(a[k & 1], *s[1:])

    def cold(k):
        s = sh[k % S]
        s[0].step(a[k & 1], *s[1:])
    count = max(S * 2, 40)
    tw, tc = graph_time(warm, count), graph_time(cold, count)
    print(json.dumps({"n_envs": n, "warm_us": tw, "cold_us": tc, "warm_frac": 284 * n / tw / 1e3 / PEAK, "cold_frac": 284 * n / tc / 1e3 / PEAK,
                      "rotating_sets": S}), flush=True)
    del sh
