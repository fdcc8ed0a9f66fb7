"""Multi-GPU metrics exchange over NVLink peer memory (csrc/peer_metrics.cu), driven from ONE process with two devices: the sum
every rank obtains equals the float64 sum of the two handles' own metrics vectors, in rank order, for several exchanges in a
row, with the push folded into a captured CUDA graph.  Skipped on a one-GPU box (the torchrun path -- cudaIpc handles exchanged
through torch.distributed -- is exercised by `bench.py --gpus N` and `benchmarks/peer_metrics_check.py`)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(dev, seed):
    from ouzelum_b200 import _lib
    from ouzelum_b200.sim import QuadSim
    n = 4096
    sim = QuadSim(_lib.default_cfg(n, seed=seed, env_id_base=seed * n, fault_mode=1, max_episode_length=9), dev)
    bufs = dict(obs=torch.zeros(n, 13, device=dev), rew=torch.zeros(n, device=dev),
                reset=torch.ones(n, dtype=torch.int64, device=dev), prog=torch.zeros(n, dtype=torch.int64, device=dev),
                tout=torch.zeros(n, dtype=torch.uint8, device=dev), epr=torch.zeros(n, device=dev))
    act = torch.rand(n, 4, device=dev, generator=torch.Generator(device=dev).manual_seed(seed)) * 2 - 1
    return sim, bufs, act


def _step(sim, b, act):
    sim.step(act, b["obs"], b["rew"], b["reset"], b["prog"], b["tout"], b["epr"])


def test_peer_metrics_two_devices_one_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from ouzelum_b200.dist import PeerMetrics
    devs = [torch.device("cuda:0"), torch.device("cuda:1")]
    sims = []
    for r, d in enumerate(devs):
        with torch.cuda.device(d):
            sims.append(_mk(d, r + 1))
    xs = [PeerMetrics(d, rank=r, world=2, connect=False) for r, d in enumerate(devs)]
    PeerMetrics.connect_same_process(xs)
    for rnd in range(12):                                   # more exchanges than ring entries
        local, summed = [], []
        for r, d in enumerate(devs):
            with torch.cuda.device(d):
                sim, b, act = sims[r]
                for _ in range(5):
                    _step(sim, b, act)
                loc = torch.zeros(16, dtype=torch.float64, device=d)
                xs[r].push(sim, local=loc)
                local.append(loc)
        for r, d in enumerate(devs):
            with torch.cuda.device(d):
                summed.append(xs[r].sum(sims[r][0]))
        for d in devs:
            torch.cuda.synchronize(d)
        want = local[0].cpu() + local[1].cpu()              # rank order
        for r in range(2):
            assert torch.equal(summed[r].cpu(), want), f"round {rnd} rank {r}"
            assert torch.equal(local[r].cpu(), sims[r][0].metrics().cpu())
        assert want[8] == 2 * 4096 * 5 * (rnd + 1) and want[9] > 0
    for x in xs:
        st = x.status()
        assert st == {"pushed": 12, "summed": 12, "error": 0}, st


def test_peer_metrics_single_rank_in_graph():
    """world = 1 degenerates to a metrics read; the push and the folded sum are graph-capturable (device-side sequence numbers)."""
    from ouzelum_b200.dist import PeerMetrics
    d = torch.device("cuda:0")
    sim, b, act = _mk(d, 3)
    x = PeerMetrics(d, rank=0, world=1)
    prev = torch.zeros(16, dtype=torch.float64, device=d)
    last = torch.zeros(16, dtype=torch.float64, device=d)
    for _ in range(3):
        _step(sim, b, act)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k in range(8):
            _step(sim, b, act)
            if k % 4 == 1:
                x.push(sim, prev_sum=prev)
        x.sum(sim, out=last)
    for rep in range(3):
        g.replay()
        torch.cuda.synchronize()
        m = sim.metrics().cpu()
        steps_total = 3 + 8 * (rep + 1)
        assert m[8] == 4096 * steps_total
        assert last.cpu()[8] == 4096 * (steps_total - 2)          # the second push of the replay sits 2 steps before its end
        assert prev.cpu()[8] == 4096 * (steps_total - 6)          # ... and folded in the sum of the first one
    assert x.status() == {"pushed": 6, "summed": 6, "error": 0}
