"""Domain randomisation with the reference's schema (isaacgymenvs/utils/dr_utils.py:71-132: gaussian / loguniform / uniform x
additive / scaling x linear / constant schedule, applied at resets -- tasks/base/vec_task.py:538-768) on the per-env body and
rotor parameters, including the motor constant (yaw_km), against both CPU oracles."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

U, LU, G = 1, 2, 3
SC, AD = 0, 1
NONE, LIN, CONST = 0, 1, 2


def _mk(n, dr, **kw):
    from ouzelum_b200 import _lib
    from ouzelum_b200.sim import QuadSim
    from oracle.quad_step import QuadStepOracle
    cfg = _lib.default_cfg(n, dr_enable=1, dr=dr, **kw)
    sim = QuadSim(cfg, DEV)
    ora = QuadStepOracle(cfg.to_dict())
    b = dict(obs=torch.zeros(n, 13, device=DEV), rew=torch.zeros(n, device=DEV), reset=torch.ones(n, dtype=torch.int64, device=DEV),
             progress=torch.zeros(n, dtype=torch.int64, device=DEV), timeout=torch.zeros(n, dtype=torch.uint8, device=DEV),
             ep_ret=torch.zeros(n, device=DEV))
    return sim, ora, b, cfg


def _params(sim):
    p, _ = sim.get_params()
    p = p.cpu()
    return torch.cat([p[:, 0:6], p[:, 7:8]], 1)             # oracle layout: mass, ixx, iyy, izz, arm, thrust scale, yaw_km


def test_uniform_schedules_and_per_env_motor_constant_bit_exact():
    """uniform draws (scaling and additive, with linear / constant schedules) are plain float32 arithmetic: bit-exact, and so is
    every step that follows -- including the per-env motor constant in the yaw torque."""
    n = 3000
    dr = {"mass": (U, SC, 0.7, 1.3, LIN, 40), "ixx": (U, AD, -0.004, 0.004, CONST, 25), "iyy": (U, SC, 0.9, 1.1),
          "izz": (U, AD, 0.0, 0.01, LIN, 1000), "arm": (U, SC, 0.95, 1.05, CONST, 10), "thrust_scale": (U, SC, 0.8, 1.2),
          "yaw_km": (U, SC, 0.5, 1.5, LIN, 30)}
    sim, ora, b, cfg = _mk(n, dr, seed=3, fault_mode=1, yaw_km=0.016, max_episode_length=18, env_id_base=77)
    from oracle.c_oracle import COracle
    cora = COracle(cfg)
    g = torch.Generator().manual_seed(1)
    seen = []
    for t in range(70):
        a = (torch.rand(n, 4, generator=g) * 2 - 1) * 0.3
        sim.step(a.to(DEV), b["obs"], b["rew"], b["reset"], b["progress"], b["timeout"], b["ep_ret"])
        ora.step(a)
        cora.step(a)
        assert torch.equal(_params(sim), ora.params), f"params t={t}"
        assert np.array_equal(_params(sim).numpy(), np.concatenate([cora.params[:, 0:6], cora.params[:, 7:8]], 1)), f"C params t={t}"
        assert torch.equal(b["obs"].cpu(), ora.obs_buf) and torch.equal(b["rew"].cpu(), ora.rew_buf), f"t={t}"
        assert torch.equal(b["reset"].cpu(), ora.reset_buf), f"t={t}"
        assert np.array_equal(b["obs"].cpu().numpy(), cora.obs_buf) and np.array_equal(b["reset"].cpu().numpy(), cora.reset_buf), f"C t={t}"
        seen.append(_params(sim)[:, 0].clone())
    p = _params(sim)
    km = p[:, 6]
    assert float(km.min()) >= 0.016 * 0.5 - 1e-9 and float(km.max()) <= 0.016 * 1.5 + 1e-9 and float(km.std()) > 1e-3
    # linear schedule: the spread of the mass scaling grows with the step at which the env was last reset
    assert bool((seen[0] == seen[0][0]).all())              # step 0: schedule scaling 0 => nominal mass everywhere
    assert float(seen[-1].std()) > 0.05


def test_gaussian_and_loguniform_draws_follow_the_oracle():
    """log-uniform and gaussian draws go through float64 log / exp / cos on both sides and are rounded to float32 once: equal to
    the oracle to 1 float32 ulp at worst (libm vs CUDA libdevice), and statistically what dr_utils.py:96-118 describes."""
    n = 20000
    dr = {"mass": (G, SC, 1.0, 0.05), "ixx": (LU, SC, 0.5, 2.0), "iyy": (G, AD, 0.0, 0.001), "izz": (LU, SC, 0.8, 1.25, LIN, 4),
          "arm": (G, SC, 1.0, 0.02, CONST, 2), "thrust_scale": (U, SC, 0.8, 1.2), "yaw_km": (G, AD, 0.0, 0.002)}
    sim, ora, b, cfg = _mk(n, dr, seed=9, yaw_km=0.01, max_episode_length=4)
    a = torch.zeros(n, 4)
    for t in range(9):
        sim.step(a.to(DEV), b["obs"], b["rew"], b["reset"], b["progress"], b["timeout"], b["ep_ret"])
        ora.step(a)
        p = _params(sim)
        torch.testing.assert_close(p, ora.params, rtol=2e-7, atol=0.0, msg=f"params t={t}")
        frac_equal = float((p == ora.params).float().mean())
        assert frac_equal > 0.999, (t, frac_equal)
        ora.params = p.clone()                              # continue from identical parameters
        assert torch.equal(b["reset"].cpu(), ora.reset_buf), f"t={t}"
        torch.testing.assert_close(b["obs"].cpu(), ora.obs_buf, rtol=1e-5, atol=1e-6)
    p = _params(sim).double()
    m0, i0 = float(cfg.mass), float(cfg.ixx)
    assert abs(float((p[:, 0] / m0).mean()) - 1.0) < 2e-3 and abs(float((p[:, 0] / m0).std()) - 0.05) < 2e-3      # N(1, 0.05)
    lr = torch.log(p[:, 1] / i0)                                                                                 # U(log .5, log 2)
    assert float(lr.min()) >= np.log(0.5) - 1e-6 and float(lr.max()) <= np.log(2.0) + 1e-6
    assert abs(float(lr.mean())) < 0.02 and abs(float(lr.std()) - (np.log(4.0) / np.sqrt(12.0))) < 0.01
    assert abs(float(p[:, 6].mean()) - 0.01) < 1e-4 and abs(float(p[:, 6].std()) - 0.002) < 1e-4                   # 0.01 + N(0, 0.002)


def test_reference_style_randomization_params_config():
    """cfg["task"]["randomize"] / ["randomization_params"] with the reference's keys (cfg/task/*.yaml) reach the kernel."""
    import ouzelum_b200
    n = 512
    cfg = ouzelum_b200.task_config("Ouzelum", n, seed=2)
    cfg["task"] = {"randomize": True, "randomization_params": {"frequency": 1, "actor_params": {"x500": {
        "rigid_body_properties": {"mass": {"range": [0.5, 1.5], "operation": "scaling", "distribution": "uniform"}},
        "rotor_properties": {"motor_constant": {"range": [0.0, 0.004], "operation": "additive", "distribution": "gaussian",
                                                "schedule": "linear", "schedule_steps": 3000}}}}}}
    env = ouzelum_b200.make(seed=2, task="Ouzelum", num_envs=n, sim_device=DEV, rl_device=DEV, headless=True, cfg=cfg)
    d = env.native_cfg.to_dict()
    assert d["dr_enable"] == 1 and d["dr"][0][:2] == (U, SC) and d["dr"][6][:2] == (G, AD) and d["dr"][6][4:] == (LIN, 3000)
    assert d["dr"][1][0] == 0                                # inertia not listed => not randomised (the reference's behaviour)
    env.step(env.zero_actions())
    p, _ = env.sim.get_params()
    assert float(p[:, 0].std()) > 0.1 and float(p[:, 1].std()) == 0.0
    bad = ouzelum_b200.task_config("Ouzelum", n)
    bad["task"] = {"randomize": True, "randomization_params": {"actor_params": {"x500": {"rigid_body_properties": {
        "mass": {"range": [0.5, 1.5], "operation": "scaling", "distribution": "cauchy"}}}}}}
    with pytest.raises(ValueError, match="distribution"):
        ouzelum_b200.make(seed=2, task="Ouzelum", num_envs=n, sim_device=DEV, rl_device=DEV, headless=True, cfg=bad)
