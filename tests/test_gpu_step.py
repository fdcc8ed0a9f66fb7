"""GPU parity of kernel K1 (`ozl_step`, through the C ABI) against the CPU oracle (oracle/quad_step.py).

float32 arithmetic is expected to agree BIT-EXACTLY (the kernel is compiled with -fmad=false and
performs the oracle's operations in the oracle's order), which makes the integer outputs (reset,
progress, timeout) exactly comparable over long roll-outs.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(n, **over):
    from ouzelum_b200 import _lib
    from ouzelum_b200.sim import QuadSim
    from oracle.quad_step import QuadStepOracle
    cfg = _lib.default_cfg(n, **over)
    sim = QuadSim(cfg, "cuda:0")
    ora = QuadStepOracle(cfg.to_dict())
    dev = torch.device("cuda:0")
    bufs = dict(
        obs=torch.zeros(n, 13, device=dev), rew=torch.zeros(n, device=dev),
        reset=torch.ones(n, dtype=torch.int64, device=dev), progress=torch.zeros(n, dtype=torch.int64, device=dev),
        timeout=torch.zeros(n, dtype=torch.uint8, device=dev), ep_ret=torch.zeros(n, device=dev))
    return sim, ora, bufs


def _compare(sim, ora, bufs, tag, exact=True):
    o_obs, o_rew, o_reset, o_to = ora.obs_buf, ora.rew_buf, ora.reset_buf, ora.timeout_buf
    assert torch.equal(bufs["reset"].cpu(), o_reset), f"{tag}: reset flags differ"
    assert torch.equal(bufs["progress"].cpu(), ora.progress_buf), f"{tag}: progress differs"
    assert torch.equal(bufs["timeout"].cpu().bool(), o_to), f"{tag}: timeout differs"
    st = sim.get_state()
    if exact:
        assert torch.equal(st["root"].cpu(), ora.root), f"{tag}: root state not bit-exact (max {(st['root'].cpu()-ora.root).abs().max()})"
        assert torch.equal(st["thrust"].cpu(), ora.thrust), f"{tag}: thrust"
        assert torch.equal(st["target"].cpu(), ora.target), f"{tag}: target"
        assert torch.equal(bufs["obs"].cpu(), o_obs), f"{tag}: obs not bit-exact (max {(bufs['obs'].cpu()-o_obs).abs().max()})"
        assert torch.equal(bufs["rew"].cpu(), o_rew), f"{tag}: reward not bit-exact"
        assert torch.equal(st["ep_ret"].cpu(), ora.ep_ret), f"{tag}: episode return"
    else:
        torch.testing.assert_close(bufs["obs"].cpu(), o_obs, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(bufs["rew"].cpu(), o_rew, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n", [1, 37, 256, 4099])
def test_step_bit_exact_random_actions(n):
    sim, ora, b = _mk(n, seed=1234)
    g = torch.Generator().manual_seed(n)
    for t in range(120):
        a = torch.rand(n, 4, generator=g) * 2 - 1
        if t % 7 == 3:
            a = a * 3.0                       # exercise clipActions
        sim.step(a.cuda(), b["obs"], b["rew"], b["reset"], b["progress"], b["timeout"], b["ep_ret"])
        ora.step(a)
        _compare(sim, ora, b, f"n={n} t={t}")
    assert sim.step_count == 120


def test_step_fault_dr_noise_bit_exact():
    n = 1000
    sim, ora, b = _mk(n, seed=7, fault_mode=1, dr_enable=1, pomdp_mode=3, pomdp_prob=0.3, noise_sigma=0.15,
                      max_episode_length=60, lin_drag=0.05, yaw_km=0.016, env_id_base=123456)
    g = torch.Generator().manual_seed(5)
    for t in range(150):
        a = (torch.rand(n, 4, generator=g) * 2 - 1) * 0.2
        sim.step(a.cuda(), b["obs"], b["rew"], b["reset"], b["progress"], b["timeout"], b["ep_ret"])
        ora.step(a)
        _compare(sim, ora, b, f"t={t}")
    p, f = sim.get_params()
    assert torch.equal(p[:, :6].cpu(), ora.params[:, :6]) and torch.equal(p[:, 7].cpu(), ora.params[:, 6])
    assert torch.equal(p[:, 6].cpu(), ora.fault_eff)
    assert torch.equal(f[:, 0].cpu().long(), ora.fault_rotor)
    assert torch.equal(f[:, 1].cpu().long(), ora.fault_onset)
    assert int(b["timeout"].sum()) >= 0
    # metrics (K6): counts exact, sums to double rounding
    m = sim.metrics().cpu().numpy()
    np.testing.assert_array_equal(m[8:16], ora.mcnt.astype(np.float64))
    np.testing.assert_allclose(m[0:2], ora.msum[0:2], rtol=1e-9)
    assert ora.mcnt[3] > 0 and ora.mcnt[6] > 0, "test should exercise time-outs and active faults"


def test_hover_stays_put_and_reward_known_answer():
    """Analytic known answer: at hover thrust the vehicle keeps its (post-reset) velocity; reward formula by hand."""
    from ouzelum_b200 import x500
    n = 64
    sim, ora, b = _mk(n)
    z = torch.zeros(n, 4, device="cuda")
    sim.step(z, b["obs"], b["rew"], b["reset"], b["progress"], b["timeout"], b["ep_ret"])
    st = sim.get_state()
    hover = float(np.float32(x500.MASS * 9.81 / 4))
    root = st["root"].clone()
    root[:, 7:13] = 0
    sim.set_state(root=root, thrust=torch.full((n, 4), hover))
    for _ in range(50):
        sim.step(z, b["obs"], b["rew"], b["reset"], b["progress"], b["timeout"], b["ep_ret"])
    st2 = sim.get_state()
    assert (st2["root"][:, 0:3] - root[:, 0:3]).abs().max() < 2e-4
    assert (st2["root"][:, 3:7] - root[:, 3:7]).abs().max() == 0
    d = (st2["target"] - st2["root"][:, 0:3]).double().norm(dim=1)
    expect = 1 / (1 + d * d) * (1 + 5.0 + 1.0)
    torch.testing.assert_close(b["rew"].double(), expect, rtol=1e-5, atol=1e-6)


def test_rollout_matches_stepping_with_same_actions():
    """Mode B (K steps in one launch, in-kernel actions) == K single steps fed the same Philox actions."""
    from oracle import philox as px
    n, K = 777, 40
    sim, ora, b = _mk(n, seed=99, fault_mode=1)
    sim2, _, b2 = _mk(n, seed=99, fault_mode=1)
    sim.rollout(K, b["obs"], b["rew"], b["reset"], b["progress"])
    ids = np.arange(n)
    for t in range(K):
        r = px.draw(99, ids, t, px.P_ACTION)
        a = torch.from_numpy(np.stack([np.float32(2.0) * px.u01(x) - np.float32(1.0) for x in r], -1))
        sim2.step(a.cuda(), b2["obs"], b2["rew"], b2["reset"], b2["progress"], b2["timeout"], b2["ep_ret"])
        ora.step(a)
    for k in ("obs", "rew", "reset", "progress"):
        assert torch.equal(b[k], b2[k]), k
    assert torch.equal(sim.get_state()["root"], sim2.get_state()["root"])
    assert torch.equal(b["obs"].cpu(), ora.obs_buf)
    assert sim.step_count == K
    np.testing.assert_array_equal(sim.metrics().cpu().numpy()[8:16], sim2.metrics().cpu().numpy()[8:16])


def test_shard_invariance():
    """Envs [a,b) of a big handle == a small handle created with env_id_base=a (multi-GPU sharding is pure slicing)."""
    n, a0 = 512, 200
    sim, _, b = _mk(n, seed=3, fault_mode=1, dr_enable=1)
    sims, _, bs = _mk(100, seed=3, fault_mode=1, dr_enable=1, env_id_base=a0)
    g = torch.Generator().manual_seed(1)
    for t in range(60):
        a = (torch.rand(n, 4, generator=g) * 2 - 1).cuda()
        sim.step(a, b["obs"], b["rew"], b["reset"], b["progress"], b["timeout"], b["ep_ret"])
        sims.step(a[a0:a0 + 100].contiguous(), bs["obs"], bs["rew"], bs["reset"], bs["progress"], bs["timeout"], bs["ep_ret"])
    assert torch.equal(b["obs"][a0:a0 + 100], bs["obs"])
    assert torch.equal(b["reset"][a0:a0 + 100], bs["reset"])
    assert torch.equal(sim.get_state()["root"][a0:a0 + 100], sims.get_state()["root"])


def test_errors_are_loud():
    from ouzelum_b200 import _lib
    from ouzelum_b200.sim import QuadSim
    with pytest.raises(RuntimeError):
        QuadSim(_lib.default_cfg(8), "cpu")
    with pytest.raises(RuntimeError):
        QuadSim(_lib.default_cfg(8, substeps=0), "cuda:0")
    with pytest.raises(RuntimeError):
        QuadSim(_lib.default_cfg(8, pomdp_mode=9), "cuda:0")


def test_kernel_vs_c_oracle_long_rollout():
    """Second, independent oracle (C restatement): 400 steps x 4096 envs with faults + DR, bit-exact."""
    from ouzelum_b200 import _lib
    from ouzelum_b200.sim import QuadSim
    from oracle.c_oracle import COracle
    n = 4096
    cfg = _lib.default_cfg(n, seed=31, fault_mode=1, dr_enable=1, max_episode_length=150)
    sim, co = QuadSim(cfg, "cuda:0"), COracle(cfg)
    dev = torch.device("cuda:0")
    b = dict(obs=torch.zeros(n, 13, device=dev), rew=torch.zeros(n, device=dev),
             reset=torch.ones(n, dtype=torch.int64, device=dev), progress=torch.zeros(n, dtype=torch.int64, device=dev),
             timeout=torch.zeros(n, dtype=torch.uint8, device=dev), ep_ret=torch.zeros(n, device=dev))
    g = torch.Generator().manual_seed(12)
    for t in range(400):
        a = (torch.rand(n, 4, generator=g) * 2 - 1) * (0.1 if t % 3 else 1.0)
        sim.step(a.cuda(), b["obs"], b["rew"], b["reset"], b["progress"], b["timeout"], b["ep_ret"])
        co.step(a.numpy())
        if t % 20 == 19 or t < 5:
            assert np.array_equal(b["reset"].cpu().numpy(), co.reset_buf), t
            assert np.array_equal(b["progress"].cpu().numpy(), co.progress_buf), t
            assert np.array_equal(b["obs"].cpu().numpy(), co.obs_buf), t
            assert np.array_equal(b["rew"].cpu().numpy(), co.rew_buf), t
    assert np.array_equal(sim.get_state()["root"].cpu().numpy(), co.root)
    assert int(co.timeout_buf.sum()) >= 0 and int(sim.metrics()[11].item()) > 0      # time-outs happened (max_len 150)


@pytest.mark.parametrize("n", [1, 129, 640, 4099])
def test_device_step_counter_record(n):
    """The device step counter (csrc/step_counter.cuh: base + (units >> shift), one unit retired per 128-env block plus the
    power-of-two padding) advances by exactly one per step launch for tile counts that are / are not powers of two,
    by K per K-step roll-out, survives set / get, and stays in step inside a replayed CUDA graph."""
    sim, _, b = _mk(n, seed=5)
    a = torch.zeros(n, 4, device="cuda")
    args = (b["obs"], b["rew"], b["reset"], b["progress"], b["timeout"], b["ep_ret"])
    for k in range(1, 6):
        sim.step(a, *args)
        assert sim.step_count == k
    sim.step_count = 10 ** 12
    assert sim.step_count == 10 ** 12
    sim.step(a, *args)
    assert sim.step_count == 10 ** 12 + 1
    sim.rollout(7, b["obs"], b["rew"], b["reset"], b["progress"])
    assert sim.step_count == 10 ** 12 + 8
    sim.step_wrench(a, None, *args)
    sim.apply_resets(b["reset"])                      # reads the counter, must not advance it
    assert sim.step_count == 10 ** 12 + 9
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        sim.step(a, *args)
        with torch.cuda.graph(g, stream=s):
            for _ in range(3):                        # odd number of launches per replay
                sim.step(a, *args)
    torch.cuda.current_stream().wait_stream(s)
    c0 = sim.step_count
    for _ in range(4):
        g.replay()
    assert sim.step_count == c0 + 12
