"""Observation / action noise lambdas of the reference's domain randomisation (tasks/base/vec_task.py:576-646).
CPU: the oracle's schedule arithmetic against a literal re-statement of the reference lines, and the moments of its samples.
GPU: `ozl_noise_lambda_apply` against the oracle on the same counter-RNG draws; `VecTask.apply_randomizations` + `step` (eager and in a
CUDA graph), the correlated component held between randomisation events and re-drawn at the next one."""
import numpy as np
import pytest
import torch

from oracle import noise_lambda as nl


def _reference_lines(p, last_step):
    """What tasks/base/vec_task.py:577-635 computes for one `observations` / `actions` block, written independently of the oracle (the
    reference needs Isaac Gym to reach these lines, so they cannot be executed here): the schedule factor s (:583-590), then for the
    `additive` operation all four numbers times s (:596-600, :623-627), for `scaling` every mean / bound blended towards 1 with s and
    every gaussian spread times s (:601-609, :628-632)."""
    kind = p.get("schedule")
    steps = p.get("schedule_steps")
    s = {None: lambda: 1, "linear": lambda: 1.0 / steps * min(last_step, steps),
         "constant": lambda: 0 if last_step < steps else 1}[kind]()
    first, second = p["range"]
    cfirst, csecond = p.get("range_correlated", [0., 0.])
    vals = [first, second, cfirst, csecond]
    if p["operation"] == "additive":
        return tuple(v * s for v in vals)
    towards_one = lambda v: v * s + 1.0 * (1.0 - s)
    if p["distribution"] == "gaussian":                     # (mu, var): the mean is blended, the spread is scaled
        return (towards_one(first), second * s, towards_one(cfirst), csecond * s)
    return tuple(towards_one(v) for v in vals)             # uniform (lo, hi): both bounds are blended


BLOCKS = [
    {"range": [0.0, 0.002], "range_correlated": [0.0, 0.001], "operation": "additive", "distribution": "gaussian",
     "schedule": "linear", "schedule_steps": 40000},                      # cfg/task/ShadowHandOpenAI_FF.yaml-style block
    {"range": [0.0, 0.05], "operation": "additive", "distribution": "gaussian"},
    {"range": [0.9, 1.1], "range_correlated": [0.95, 1.05], "operation": "scaling", "distribution": "uniform",
     "schedule": "constant", "schedule_steps": 100},
    {"range": [1.0, 0.1], "range_correlated": [1.0, 0.02], "operation": "scaling", "distribution": "gaussian",
     "schedule": "linear", "schedule_steps": 10},
    {"range": [-0.1, 0.3], "operation": "additive", "distribution": "uniform", "schedule": "linear", "schedule_steps": 8},
]


@pytest.mark.parametrize("p", BLOCKS)
@pytest.mark.parametrize("last_step", [0, 3, 99, 100, 50000])
def test_schedule_arithmetic_equals_the_reference_lines(p, last_step):
    got = nl.scheduled_params(p, last_step)
    want = _reference_lines(p, last_step)
    assert got[0] == (nl.GAUSSIAN if p["distribution"] == "gaussian" else nl.UNIFORM)
    assert got[1] == (nl.ADDITIVE if p["operation"] == "additive" else nl.SCALING)
    np.testing.assert_array_equal(np.array(got[2:], dtype=np.float32), np.array(want, dtype=np.float32))


def test_oracle_samples_have_the_moments_the_reference_formula_implies():
    n, w = 200000, 4
    x = np.ones((n, w), dtype=np.float32) * 2.0
    spec = nl.scheduled_params({"range": [0.5, 0.2], "range_correlated": [0.1, 0.3], "operation": "additive", "distribution": "gaussian"}, 0)
    y = nl.noise_lambda(x, spec, seed=3, step=7, corr_epoch=0)
    d = (y - x).astype(np.float64)
    assert abs(d.mean() - 0.6) < 3e-3 and abs(d.std() - np.hypot(0.2, 0.3)) < 3e-3       # N(0.5, 0.2) + N(0.1, 0.3)
    y2 = nl.noise_lambda(x, spec, seed=3, step=8, corr_epoch=0)                          # next step, same event: corr is held
    c = np.corrcoef((y - x).ravel(), (y2 - x).ravel())[0, 1]
    assert abs(c - 0.09 / 0.13) < 1e-2                                                  # var_corr^2 / (var^2 + var_corr^2)
    y3 = nl.noise_lambda(x, spec, seed=3, step=8, corr_epoch=8)                          # next event: corr re-drawn
    assert abs(np.corrcoef((y - x).ravel(), (y3 - x).ravel())[0, 1]) < 1e-2
    specu = nl.scheduled_params({"range": [0.9, 1.1], "operation": "scaling", "distribution": "uniform"}, 0)
    yu = nl.noise_lambda(x, specu, seed=3, step=7, corr_epoch=0)
    r = (yu / x).astype(np.float64)
    assert r.min() >= 0.9 - 1e-6 and r.max() < 1.1 + 1e-6 and abs(r.mean() - 1.0) < 1e-3
    assert abs(np.corrcoef(d[:, 0], d[:, 1])[0, 1]) < 1e-2                               # elements are independent


@pytest.mark.gpu
@pytest.mark.parametrize("p", BLOCKS)
@pytest.mark.parametrize("which,width", [(0, 13), (1, 4), (0, 21)])
def test_kernel_equals_oracle(p, which, width):
    import ctypes as C
    from ouzelum_b200._lib import OzlNoiseLambda, check, lib
    n, seed, step, epoch, base, clip = 3001, 11, 12345, 12000, 77, 0.9
    g = torch.Generator().manual_seed(width)
    x = (torch.rand(n, width, generator=g) * 2 - 1)
    spec = nl.scheduled_params(p, 37)
    s = OzlNoiseLambda(spec[0], spec[1], float(spec[2]), float(spec[3]), float(spec[4]), float(spec[5]))
    xd = x.cuda()
    check(lib.ozl_noise_lambda_apply(n, width, xd.data_ptr(), C.byref(s), clip, seed, step, None, 0, epoch, base, which,
                                     torch.cuda.current_stream().cuda_stream))
    want = nl.noise_lambda(x.numpy(), spec, seed, step, epoch, env_id_base=base, which=which, clip=clip)
    got = xd.cpu().numpy()
    # float64 log / sincos on the two sides may differ by an ulp before the one rounding to float32
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-7)
    assert (got == want).mean() > 0.99


@pytest.mark.gpu
def test_vectask_applies_the_lambdas_like_the_reference_step():
    import ouzelum_b200
    from oracle.quad_step import QuadStepOracle
    n, seed = 512, 21
    mk = lambda: ouzelum_b200.make(seed=seed, task="Ouzelum", num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True,
                                   cfg=ouzelum_b200.task_config("Ouzelum", n, seed=seed))
    env, plain = mk(), mk()
    dr = {"frequency": 5,
          "observations": {"range": [0.0, 0.02], "range_correlated": [0.0, 0.01], "operation": "additive", "distribution": "gaussian"},
          "actions": {"range": [0.95, 1.05], "operation": "scaling", "distribution": "uniform", "schedule": "linear", "schedule_steps": 4}}
    env.apply_randomizations(dr)
    assert env.dr_randomizations["observations"]["var"] == 0.02 and env.dr_randomizations["actions"]["lo"] == 1.0   # schedule at step 0
    g = torch.Generator().manual_seed(0)
    clip = float(env.clip_obs)
    for t in range(12):
        a = torch.rand(n, 4, generator=g) * 2 - 1
        if t in (3, 7):                                   # 7 - 0 >= 5: a new randomisation event; 3 - 0 < 5: nothing changes
            before = dict(env.dr_randomizations["actions"])
            env.apply_randomizations(dr)
            after = env.dr_randomizations["actions"]
            assert (after["corr_epoch"] == 7 and abs(after["lo"] - 0.95) < 1e-7) if t == 7 else (after == before)
        spec_a = nl.scheduled_params(dr["actions"], env.dr_randomizations["actions"]["corr_epoch"])
        spec_o = nl.scheduled_params(dr["observations"], env.dr_randomizations["observations"]["corr_epoch"])
        ep = env.dr_randomizations["actions"]["corr_epoch"]
        a_noised = nl.noise_lambda(a.numpy(), spec_a, seed, t, ep, which=1)
        o, r, d, _ = env.step(a.cuda())
        po, pr, pd, _ = plain.step(torch.from_numpy(a_noised).cuda())    # the same env fed the noised actions, no lambdas
        assert torch.equal(r, pr) and torch.equal(d, pd), t              # the dynamics saw the noised actions
        want = nl.noise_lambda(po["obs"].cpu().numpy(), spec_o, seed, t, ep, which=0, clip=clip)
        np.testing.assert_allclose(o["obs"].cpu().numpy(), want, rtol=2e-6, atol=2e-7, err_msg=f"t={t}")
    assert env.sim.step_count == 12


@pytest.mark.gpu
def test_lambdas_are_graph_capturable():
    import ouzelum_b200
    n = 256
    mk = lambda: ouzelum_b200.make(seed=2, task="Ouzelum", num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True,
                                   cfg=ouzelum_b200.task_config("Ouzelum", n, seed=2))
    dr = {"observations": {"range": [0.0, 0.03], "operation": "additive", "distribution": "gaussian"},
          "actions": {"range": [0.0, 0.1], "operation": "additive", "distribution": "uniform"}}
    e1, e2 = mk(), mk()
    e1.apply_randomizations(dr), e2.apply_randomizations(dr)
    a = torch.rand(n, 4, device="cuda:0") * 2 - 1
    for e in (e1, e2):
        e.step(a)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(4):
            e2.step(a)
    for _ in range(2):
        gr.replay()
        for _ in range(4):
            e1.step(a)
        torch.cuda.synchronize()
        assert torch.equal(e1.obs_buf, e2.obs_buf) and torch.equal(e1.rew_buf, e2.rew_buf)
