"""Direct oracle parity of the ONE-LAUNCH EKFLeeLanded control step (`ozl_ekf_lee_landed_step`, the kernel BASELINE config 3 is
quoted on) at config-3 scale: 65536 envs, domain randomisation on, sensor noise sigma 0.15, attitude EKF + PV filter + Lee
controller in the loop.  Reference: isaacgymenvs/tasks/ekf_lee_landed.py:308-530 (pre_physics_step) and :620-665.

Every step is compared from IDENTICAL state: the device state (env planes, per-env parameters, filter banks, glue buffers) is
copied into the oracles before the step, the kernel and the oracles then advance once.
  * estimator / controller: oracle/ekf_lee_landed.py (EKFLeeGlue) -- EKF <= 1e-9, PV / waypoint / wrench at the stated tolerances
  * physics / observation / reward / reset: oracle/quad_step.py (QuadStepOracle, wrench actuation) fed with the KERNEL's wrench --
    the step arithmetic in this kernel is the same explicit-FMA code as in quad_step.cu, so given the same wrench every output is
    compared BIT FOR BIT (ints and floats)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _snapshot_into(env, phys, glue, t):
    st = env.sim.get_state()
    params, fault = env.sim.get_params()
    phys.load(st["root"].cpu().numpy(), st["thrust"].cpu().numpy(), st["target"].cpu().numpy(), st["ep_ret"].cpu().numpy(),
              params.cpu().numpy(), fault.cpu().numpy(), env.reset_buf.cpu().numpy(), env.progress_buf.cpu().numpy(), t)
    glue.Q = env.ekf.Q_state.cpu().numpy().copy()
    glue.ekf.P = env.ekf.P.cpu().numpy().copy()
    glue.pv.state = env.pvfilters.get_states().cpu().numpy().copy()
    glue.pv.cov = env.pvfilters.get_covariances().cpu().numpy().copy()
    glue.prev_v = env.prev_root_linvels.cpu().numpy().copy()
    glue.waypoints = env.target_waypoints.cpu().numpy().copy()
    glue.step = t


@pytest.mark.parametrize("n,steps", [(65536, 9), (4099, 24)])
def test_one_launch_ekf_lee_landed_step_vs_oracle(n, steps):
    import ouzelum_b200
    from oracle.ekf_lee_landed import EKFLeeGlue
    from oracle.lee_control import lee_control
    from oracle.quad_step import QuadStepOracle
    conv, seed, sigma = 3, 12, 0.15
    cfg = ouzelum_b200.task_config("EKFLeeLanded", n, seed=seed, POMDP="random_noise", pomdp_prob=sigma, ConvergenceTime=conv,
                                   domainRandomization={"enable": True}, rotorFault={"enable": True}, maxEpisodeLength=7,
                                   exposeEstimates=True)        # the kernel also writes its estimate / command (`_est`, `_cmd`)
    env = ouzelum_b200.make(seed=seed, task="EKFLeeLanded", num_envs=n, sim_device=DEV, rl_device=DEV, headless=True, cfg=cfg)
    assert env.fused and env.fused_step                       # the one-launch path
    phys = QuadStepOracle(env.native_cfg.to_dict())
    assert phys.cfg["wrench_warmup_steps"] == conv and phys.cfg["dr_enable"] == 1 and phys.cfg["pomdp_mode"] == 2
    glue = EKFLeeGlue(n, convergence=conv, pomdp_mode=2, pomdp_prob=sigma, seed=seed)
    a = torch.zeros(n, 4, device=DEV)
    n_reset = n_fix = 0
    for t in range(steps):
        _snapshot_into(env, phys, glue, t)
        reset_before = env.reset_buf.bool().cpu().numpy().copy()
        n_reset += int(reset_before.sum())
        env.step(a)
        warm = t < conv
        tgt = env.husky.target.cpu().numpy()                # the vehicle kernel is pinned separately (test_gpu_tasks.py, K5)
        root = phys.post_reset_root().numpy()               # what the estimator chain saw: state after reset_idx (:312-314)
        wrench_o, est_o, cmd_o = glue.pre_physics(root, tgt, reset_before)
        # ---- estimator / controller (E1-E3, V1-V2, L, L')
        np.testing.assert_allclose(env.ekf.Q_state.cpu().numpy(), glue.Q, rtol=1e-9, atol=1e-12, err_msg=f"EKF q t={t}")
        np.testing.assert_allclose(env.ekf.P.cpu().numpy(), glue.ekf.P, rtol=1e-7, atol=1e-9 * np.abs(glue.ekf.P).max(), err_msg=f"EKF P t={t}")
        assert np.array_equal(env.prev_root_linvels.cpu().numpy(), root[:, 7:10])
        x, P = env.pvfilters.get_states().cpu().numpy(), env.pvfilters.get_covariances().cpu().numpy()
        sx, sP = np.abs(glue.pv.state).max() + 1.0, np.abs(glue.pv.cov).max() + 1.0
        # float32 Kalman update with P ~ 1e3, R = 1e-7 (and R = 0 in the velocity fix): see test_pv_filter_float32_error_vs_float64
        np.testing.assert_allclose(x, glue.pv.state, rtol=2e-3, atol=2e-4 * sx, err_msg=f"PV x t={t}")
        tight = np.isclose(x, glue.pv.state, rtol=1e-4, atol=2e-5 * sx)
        assert tight.mean() > 0.999, f"PV x t={t}: {(~tight).sum()} of {tight.size} state entries beyond 1e-4"
        tightP = np.isclose(P, glue.pv.cov, rtol=1e-3, atol=2e-6 * sP)
        assert tightP.mean() > 0.999, f"PV P t={t}: {(~tightP).sum()} covariance entries off"
        np.testing.assert_allclose(P, glue.pv.cov, rtol=2e-2, atol=3e-4 * sP, err_msg=f"PV P t={t}")
        k = t * n + np.arange(n)
        n_fix += int(((k % 7) == 6).sum() + ((k % 3) == 0).sum())
        np.testing.assert_allclose(env._cmd.cpu().numpy(), cmd_o, rtol=1e-5, atol=1e-5, err_msg=f"cmd t={t}")
        np.testing.assert_allclose(env.target_waypoints.cpu().numpy(), glue.waypoints, rtol=1e-5, atol=1e-5)
        w = env._wrench.cpu().numpy()
        if warm:
            assert np.array_equal(w, wrench_o), f"hover wrench t={t}"
        else:
            np.testing.assert_allclose(env._est.cpu().numpy(), est_o, rtol=2e-3, atol=2e-4 * sx, err_msg=f"est t={t}")
            # the controller amplifies estimate differences by its gains: compare it on the kernel's own estimate
            th, tq = lee_control(env._est.cpu().numpy(), env._cmd.cpu().numpy(), mode=0)
            np.testing.assert_allclose(w[:, 0], glue.mg * th, rtol=2e-5, atol=2e-4, err_msg=f"thrust t={t}")
            np.testing.assert_allclose(w[:, 1:], tq, rtol=2e-5, atol=1e-4, err_msg=f"torque t={t}")
        # ---- physics + observation + reward + reset, given the kernel's wrench: bit for bit
        obs_o, rew_o, reset_o, tout_o = phys.step(torch.from_numpy(w), target_in=torch.from_numpy(tgt), act_mode=1)
        assert torch.equal(env.reset_buf.cpu(), reset_o), f"reset t={t}: {(env.reset_buf.cpu() != reset_o).sum()} differ"
        assert torch.equal(env.progress_buf.cpu(), phys.progress_buf), f"progress t={t}"
        assert torch.equal(env.timeout_buf.cpu(), tout_o), f"timeout t={t}"
        assert torch.equal(env.obs_buf.cpu(), obs_o), f"obs t={t} max {(env.obs_buf.cpu() - obs_o).abs().max()}"
        assert torch.equal(env.rew_buf.cpu(), rew_o), f"reward t={t}"
        st = env.sim.get_state()
        assert torch.equal(st["root"].cpu(), phys.root), f"root t={t}"
        assert torch.equal(st["ep_ret"].cpu(), phys.ep_ret)
        params, fault = env.sim.get_params()
        assert torch.equal(params[:, :6].cpu(), phys.params[:, :6]), f"DR params t={t}"
        assert torch.equal((fault[:, 1] < 0).cpu(), phys.landed), f"landed flag t={t}"
    assert n_reset > n and n_fix > 0 and env.episodes > 0      # resets (with DR draws) and sensor fixes were exercised


def test_shared_trigger_counters_are_shard_invariant():
    """ekf_lee_landed.py:425-440: the position / velocity trigger counters are shared by all envs and advance once per
    env-iteration, so which env gets a fix at step t depends on the TOTAL env count.  Two half-size shards with
    numEnvsTotal = N must reproduce the single-handle run bit for bit (the index was built from the local env id)."""
    import ouzelum_b200
    n, h = 1024, 512
    mk = lambda num, base, total: ouzelum_b200.make(
        seed=5, task="EKFLeeLanded", num_envs=num, sim_device=DEV, rl_device=DEV, headless=True,
        cfg=ouzelum_b200.task_config("EKFLeeLanded", num, seed=5, POMDP="random_noise", pomdp_prob=0.05, ConvergenceTime=4,
                                     maxEpisodeLength=20, envIdBase=base, numEnvsTotal=total))
    full, lo, hi = mk(n, 0, n), mk(h, 0, n), mk(h, h, n)
    a, ah = torch.zeros(n, 4, device=DEV), torch.zeros(h, 4, device=DEV)
    for t in range(30):
        o, r, d, _ = full.step(a)
        o1, r1, d1, _ = lo.step(ah)
        o2, r2, d2, _ = hi.step(ah)
        assert torch.equal(o["obs"], torch.cat([o1["obs"], o2["obs"]])), t
        assert torch.equal(r, torch.cat([r1, r2])) and torch.equal(d, torch.cat([d1, d2])), t
    assert torch.equal(full.pvfilters.get_states(), torch.cat([lo.pvfilters.get_states(), hi.pvfilters.get_states()]))


@pytest.mark.parametrize("n,force", [(65536, True), (4099, True), (300, True), (100000, False)])
def test_tile_chained_graph_replay_equals_eager_steps(n, force, monkeypatch):
    """Steps captured back to back into one CUDA graph are launched TILE-CHAINED (csrc/tile_chain.cuh: launch L+1's CTA b waits only
    for launch L's CTA b, not for the whole grid, so consecutive launches overlap).  The replay must leave every buffer exactly as
    the same number of eager (classic, fully ordered) launches does: env planes, filter banks, glue buffers, surface tensors,
    episode statistics and the step counter -- over several replays, with eager steps in between (the first launch of every
    replay re-derives the tile sequence from the global step record)."""
    import ouzelum_b200
    if force:
        monkeypatch.setenv("OZL_EKF_CHAIN", "2")  # chain one-wave grids too (the default chains only grids longer than one wave)
    else:
        monkeypatch.delenv("OZL_EKF_CHAIN", raising=False)     # 100000 envs = 1042 CTAs > 740 resident: chained by default
    mk = lambda: ouzelum_b200.make(
        seed=3, task="EKFLeeLanded", num_envs=n, sim_device=DEV, rl_device=DEV, headless=True,
        cfg=ouzelum_b200.task_config("EKFLeeLanded", n, seed=3, POMDP="random_noise", pomdp_prob=0.15, ConvergenceTime=5,
                                     domainRandomization={"enable": True}, rotorFault={"enable": True}, maxEpisodeLength=11))
    eager, graphed = mk(), mk()
    a = torch.zeros(n, 4, device=DEV)
    K = 17
    for e in (eager, graphed):
        for _ in range(3):
            e.step(a)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(K):
            graphed._launch(a)
    torch.cuda.synchronize()
    # the capture itself launched nothing: bring the host-side step count back
    graphed.sim_step_count -= K

    def same(tag):
        torch.cuda.synchronize()
        se, sg = eager.sim.get_state(), graphed.sim.get_state()
        for k_ in se:
            assert torch.equal(se[k_], sg[k_]), f"{tag}: sim state {k_}"
        pe, pg = eager.sim.get_params(), graphed.sim.get_params()
        assert torch.equal(pe[0], pg[0]) and torch.equal(pe[1], pg[1]), f"{tag}: params"
        for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "prev_root_linvels", "target_waypoints", "_wrench"):
            assert torch.equal(getattr(eager, name), getattr(graphed, name)), f"{tag}: {name}"
        assert torch.equal(eager.ekf._q, graphed.ekf._q) and torch.equal(eager.ekf._P, graphed.ekf._P), f"{tag}: EKF"
        assert torch.equal(eager.pvfilters._x, graphed.pvfilters._x), f"{tag}: PV state"
        assert torch.equal(eager.pvfilters._P, graphed.pvfilters._P), f"{tag}: PV covariance"
        assert torch.equal(eager.husky.target, graphed.husky.target), f"{tag}: vehicle"
        me, mg = eager.sim.metrics().cpu(), graphed.sim.metrics().cpu()
        assert torch.equal(me[[2, 8, 9, 10, 11, 12, 13, 14, 15]], mg[[2, 8, 9, 10, 11, 12, 13, 14, 15]]), f"{tag}: integer metrics"
        assert eager.sim.step_count == graphed.sim.step_count, f"{tag}: step counter"

    for rep in range(3):
        g.replay()
        graphed.sim_step_count += K
        for _ in range(K):
            eager._launch(a)
        same(f"replay {rep}")
        for e in (eager, graphed):            # eager steps between replays
            e.step(a)
            e.step(a)
        same(f"eager after replay {rep}")
    assert eager.episodes > 0
