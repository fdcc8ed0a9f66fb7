"""GPU parity of the companion kernels (Lee controllers K4, PV filter K3, attitude EKF K2, sensor-fault K7, episode
statistics) through the C ABI / Python mirrors, against (a) the golden fixtures produced by the reference's own code and
(b) the CPU oracle on seeded random inputs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


# ------------------------------------------------------------------------------------------------ K4
@pytest.mark.parametrize("name", ["lee_position_control", "lee_velocity_control", "lee_attitude_control"])
def test_lee_controller_vs_reference_fixture_and_oracle(name):
    from ouzelum_b200.controllers import Controller, control
    from oracle.lee_control import lee_control
    d = load("lee.npz")
    cc = control()
    cc.controller = name
    c = Controller(cc, DEV)
    st, cmd = torch.from_numpy(d["state"]).to(DEV), torch.from_numpy(d["command"]).to(DEV)
    thrust, torque = c(st, cmd)
    # float32 with sin/cos/atan2/asin: 1e-5 against the reference's output wherever the state is away from the
    # singular sets of the Euler extraction (|pitch| -> pi/2: asin' blows up, yaw / roll atan2 degenerate); everywhere else the
    # looser bound of round 1 holds
    th, tq = thrust.cpu().numpy(), torque.cpu().numpy()
    q = d["state"][:, 3:7].astype(np.float64)
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    sin_pitch = 2.0 * (q[:, 3] * q[:, 1] - q[:, 2] * q[:, 0])               # -R[2][0]
    regular = np.abs(sin_pitch) < 0.95
    assert regular.mean() > 0.6
    np.testing.assert_allclose(th[regular], d[name + "_thrust"][regular], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(tq[regular], d[name + "_torque"][regular], rtol=1e-5, atol=1e-5 * max(1.0, float(np.abs(d[name + "_torque"]).max()) / 10))
    np.testing.assert_allclose(th, d[name + "_thrust"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(tq, d[name + "_torque"], rtol=2e-5, atol=5e-5)
    err = np.abs(tq - d[name + "_torque"])
    print(name, "max |torque err| regular:", float(err[regular].max()), "all:", float(err.max()), "max |torque|:", float(np.abs(tq).max()))
    t64, q64 = lee_control(d["state"], d["command"], mode={"lee_position_control": 0, "lee_velocity_control": 1,
                                                           "lee_attitude_control": 2}[name], dtype=np.float64)
    np.testing.assert_allclose(thrust.cpu().numpy(), t64, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(torque.cpu().numpy(), q64, rtol=2e-5, atol=5e-5)


def test_lee_known_answers_and_errors():
    from ouzelum_b200.controllers import Controller, control
    d = load("lee.npz")
    c = Controller(control(), DEV)
    thrust, torque = c(torch.from_numpy(d["kat_state"]).to(DEV), torch.tensor([[0.0, 0, 1, 0]] * 2, device=DEV))
    np.testing.assert_allclose(thrust.cpu().numpy(), [1.0, 2.0385], atol=1e-4)
    np.testing.assert_allclose(torque.cpu().numpy(), [[0, 0, 0], [-0.5959, -0.1204, 0.0262]], atol=1e-4)
    bad = control()
    bad.controller = "pid"
    with pytest.raises(ValueError, match="Invalid controller name"):
        Controller(bad, DEV)
    with pytest.raises(RuntimeError):
        Controller(control(), "cpu")


# ------------------------------------------------------------------------------------------------ K3
def test_pv_filter_vs_reference_fixture():
    from ouzelum_b200.pv_filter import PVFilterBank
    d = load("pvfilter.npz")
    T, n = d["acc"].shape[:2]
    bank = PVFilterBank(n, [1.0, 1.0, 1.0], DEV)
    var = [1e-7] * 3
    t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    for t in range(T):
        # one fused launch: predict + gated position fix + gated velocity fix (reference order, ekf_lee_landed.py:419-440)
        bank.step(accels=t_(d["acc"][t]), orientation=t_(d["quat"][t]), dt=0.01, flip_Qw=(t % 2 == 0),
                  gps_data=t_(d["pos_meas"][t]), gps_var=var, gps_mask=t_(d["pos_fix"][t]),
                  vel_data=t_(d["vel_meas"][t]), vel_var=None, vel_mask=t_(d["vel_fix"][t]))      # R = 0 in the velocity fix: the
        # fixture was made the way the task calls the filter -- correction_step(vel_data, vel_var) with gps_var=None (PVFilter.py:76-79)
        x = bank.get_states().cpu().numpy()
        scale = np.abs(d["states"][t]).max() + 1.0
        np.testing.assert_allclose(x, d["states"][t], rtol=1e-4, atol=1e-4 * scale, err_msg=f"t={t}")
        bank.set_states(t_(d["states"][t]))
        bank.set_covariances(t_(d["covs"][t]))


def test_pv_filter_float32_error_vs_float64_arbiter():
    """V2 pin on the GPU (see tests/test_oracle_golden.py::test_pv_filter_float32_error_is_inherent_float64_arbiter): from the
    reference's own previous state, the kernel's float32 step must be as close to the float64 evaluation as the reference's
    float32 step is (state and covariance), and must match the reference to 1e-5 where only the prediction runs."""
    from ouzelum_b200.pv_filter import PVFilterBank
    from test_oracle_golden import _pv_single_step, pv_fixture_steps
    d = load("pvfilter.npz")
    n = d["acc"].shape[1]
    bank = PVFilterBank(n, [1.0, 1.0, 1.0], DEV)
    var = [1e-7] * 3
    t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    ratios = []
    for t, ps, pc in pv_fixture_steps(d):
        bank.set_states(t_(ps))
        bank.set_covariances(t_(pc))
        bank.step(accels=t_(d["acc"][t]), orientation=t_(d["quat"][t]), dt=0.01, flip_Qw=(t % 2 == 0),
                  gps_data=t_(d["pos_meas"][t]), gps_var=var, gps_mask=t_(d["pos_fix"][t]),
                  vel_data=t_(d["vel_meas"][t]), vel_var=None, vel_mask=t_(d["vel_fix"][t]))      # R = 0 in the velocity fix: the
        # fixture was made the way the task calls the filter -- correction_step(vel_data, vel_var) with gps_var=None (PVFilter.py:76-79)
        x, P = bank.get_states().cpu().numpy(), bank.get_covariances().cpu().numpy()
        s64, c64 = _pv_single_step(d, t, ps, pc, np.float64)
        sref, cref = d["states"][t], d["covs"][t]
        sc, scc = np.abs(s64).max() + 1.0, np.abs(c64).max()
        e_ref_s, e_gpu_s = np.abs(sref - s64).max() / sc, np.abs(x - s64).max() / sc
        e_ref_c, e_gpu_c = np.abs(cref - c64).max() / scc, np.abs(P - c64).max() / scc
        assert e_gpu_s <= 4.0 * e_ref_s + 1e-6, (t, e_gpu_s, e_ref_s)       # floor: 8 float32 ulp of the scale
        assert e_gpu_c <= 4.0 * e_ref_c + 1e-6, (t, e_gpu_c, e_ref_c)
        ratios.append((e_gpu_s / max(e_ref_s, 1e-12), e_gpu_c / max(e_ref_c, 1e-12)))
        nofix = ~(d["pos_fix"][t] | d["vel_fix"][t])
        np.testing.assert_allclose(x[nofix], sref[nofix], rtol=1e-5, atol=1e-5 * sc, err_msg=f"predict-only state t={t}")
        np.testing.assert_allclose(P[nofix], cref[nofix], rtol=1e-5, atol=1e-5 * scc, err_msg=f"predict-only cov t={t}")
        np.testing.assert_allclose(P, cref, rtol=0, atol=4.0 * e_ref_c * scc + 1e-6 * scc, err_msg=f"cov t={t}")
    print("GPU / reference float32 error ratios against float64 (state, cov) per step:", [(round(a, 2), round(b, 2)) for a, b in ratios])


def test_pv_filter_vs_oracle_single_steps_and_trigger_rule():
    from ouzelum_b200.pv_filter import PVFilterBank
    from oracle.pv_filter import PVFilterBank as Ora
    n = 1000
    rng = np.random.default_rng(3)
    bank, ora = PVFilterBank(n, [1.0, 1.0, 1.0], DEV), Ora(n, [1.0, 1.0, 1.0])
    t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    for t in range(6):
        acc = rng.normal(size=(n, 3)).astype(np.float32) * 3
        q = rng.normal(size=(n, 4)).astype(np.float32)
        pm, vm = rng.normal(size=(n, 3)).astype(np.float32), rng.normal(size=(n, 3)).astype(np.float32)
        k = t * n + np.arange(n)
        pos_fix, vel_fix = (k % 7) == 6, (k % 3) == 0
        bank.step(accels=t_(acc), orientation=t_(q), dt=0.01, flip_Qw=False, gps_data=t_(pm), gps_var=[1e-7] * 3,
                  vel_data=t_(vm), vel_var=[1e-7] * 3, trigger=(7, 6, 3, 0), iter_base=t * n)
        ora.prediction_step(acc, q, 0.01, flip_Qw=False)
        ora.correction_step(gps_data=pm, gps_var=np.full(3, 1e-7, np.float32), mask=pos_fix)
        ora.correction_step(vel_data=vm, vel_var=np.full(3, 1e-7, np.float32), gps_var=np.full(3, 1e-7, np.float32), mask=vel_fix)
        x, P = bank.get_states().cpu().numpy(), bank.get_covariances().cpu().numpy()
        sx, sP = np.abs(ora.state).max() + 1, np.abs(ora.cov).max() + 1
        np.testing.assert_allclose(x, ora.state, rtol=1e-3, atol=1e-4 * sx, err_msg=f"t={t}")
        # float32 Kalman update with P ~ 1e3 and R = 1e-7 is ill-conditioned (catastrophic cancellation in (I-KH)P): the bulk
        # of the covariance agrees to 1e-3, a handful of entries only to ~1e-4 of the covariance scale -- in the reference too
        tight = np.isclose(P, ora.cov, rtol=1e-3, atol=2e-6 * sP)
        assert tight.mean() > 0.999, f"t={t}: {(~tight).sum()} covariance entries off"
        np.testing.assert_allclose(P, ora.cov, rtol=2e-2, atol=3e-4 * sP, err_msg=f"t={t}")
        ora.state, ora.cov = x.copy(), P.copy()                               # identical state for the next single step
    # known answer (SURVEY KAT-V)
    b = PVFilterBank(1, [1.0, 1.0, 1.0], DEV)
    b.prediction_step(torch.tensor([[0.0, 0, 9.8]]), torch.tensor([[0.0, 0, 0, 1]]), dt=0.01)
    np.testing.assert_allclose(b.get_states().cpu().numpy()[0], [0, 0, 4.9e-4, 0, 0, 0.098, 0, 0, 0], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(np.diag(b.get_covariances().cpu().numpy()[0]), [1000.1] * 6 + [1000] * 3, rtol=1e-4)
    b.correction_step(gps_data=torch.tensor([[1.0, 2, 3]]), gps_var=[1e-7] * 3)
    np.testing.assert_allclose(b.get_states().cpu().numpy()[0][:6], [1, 2, 3, 0.0099995, 0.019999, 0.12799], rtol=2e-3, atol=1e-5)


# ------------------------------------------------------------------------------------------------ K2
def test_ekf_vs_reference_fixture_and_oracle():
    from ouzelum_b200.ahrs_ekf import EKFBank
    from oracle.ahrs_ekf import EKFBank as Ora
    d = load("ekf.npz")
    T, n = d["gyr"].shape[:2]
    bank = EKFBank(n, frequency=100.0, device=DEV)
    bank.set_state(torch.from_numpy(d["q0"]))
    # the fixture's sensor streams are float64; the kernel takes the float32 sensor tensors the tasks have
    # (tasks/ekf_lee_landed.py:370-381 converts float32 tensors to numpy) -> feed float32-rounded values to both
    ora = Ora(n, 100.0)
    q = d["q0"].copy()
    for t in range(T):
        g32, a32 = d["gyr"][t].astype(np.float32), d["ang"][t].astype(np.float32)
        Q = bank.update(torch.from_numpy(g32).to(DEV), torch.from_numpy(a32).to(DEV)).cpu().numpy()
        q = ora.update(q / np.linalg.norm(q, axis=1, keepdims=True), g32.astype(np.float64), a32.astype(np.float64))
        np.testing.assert_allclose(Q, q, rtol=1e-9, atol=1e-12, err_msg=f"t={t}")          # float64 kernel vs float64 oracle
        np.testing.assert_allclose(bank.P.cpu().numpy(), ora.P, rtol=1e-7, atol=1e-15)
        np.testing.assert_allclose(Q, d["Q"][t], rtol=0, atol=2e-7)                        # vs reference run on float64 sensors
    b = EKFBank(1, device=DEV)
    kq = b.update(torch.tensor([[0.1, -0.2, 0.3]]), torch.tensor([[0.9990, 0.03, -0.02, 0.025]])).cpu().numpy()[0]
    np.testing.assert_allclose(kq, [0.99903697, 0.03000111, -0.02000074, 0.02500092], atol=1e-7)   # SURVEY KAT-E
    b.set_state(torch.tensor([[2.0, 0, 0, 0]]))
    with pytest.raises(ValueError):
        b.update(torch.zeros(1, 3), torch.tensor([[1.0, 0, 0, 0]]), check_norm=True)


# ------------------------------------------------------------------------------------------------ K7 + statistics
def test_pomdp_wrapper_matches_in_kernel_noise_and_oracle():
    from ouzelum_b200.pomdp import POMDPWrapper
    from oracle import philox as px
    n = 513
    obs = torch.randn(n, 13, device=DEV)
    w = POMDPWrapper("random_noise", 0.15, seed=5, env_id_base=100, stream_id=0)
    out = w.observation(obs)
    us = []
    for k in range(4):
        us += list(px.draw(5, np.arange(n) + 100, 0, px.P_OBSNOISE + k))
    lo = np.float32(1.0 - np.float32(0.15).astype(np.float64))
    rng = np.float32(1.0 + np.float32(0.15).astype(np.float64)) - lo
    noise = np.stack([px.u01(us[j]) * rng + lo for j in range(13)], -1)
    assert np.array_equal(out.cpu().numpy(), obs.cpu().numpy() * noise)
    assert (noise >= 0.85 - 1e-6).all() and (noise <= 1.15 + 1e-6).all()
    # flicker: whole-batch blackout with probability p, decided by ONE draw per call (POMDP.py:25)
    f = POMDPWrapper("flicker", 0.3, seed=9, stream_id=0)
    blk = []
    for call in range(200):
        o = f.observation(obs)
        z = bool((o == 0).all())
        assert z or torch.equal(o, obs)
        expect = px.u01(px.draw(9, np.array([px.GLOBAL_ENV]), call, px.P_FLICKER)[0])[0] <= np.float32(0.3)
        assert z == bool(expect)
        blk.append(z)
    assert 0.15 < np.mean(blk) < 0.45
    with pytest.raises(ValueError):
        POMDPWrapper("random_sensor_missing", 0.1)


def test_episode_statistics_wrapper_matches_reference_formula():
    from ouzelum_b200.wrappers import RecordEpisodeStatisticsTorch

    class Fake:
        num_envs = 300

        def reset(self):
            return {"obs": None}

        def step(self, a):
            return {"obs": None}, a[0], a[1], {}
    w = RecordEpisodeStatisticsTorch(Fake(), DEV)
    w.reset()
    g = torch.Generator().manual_seed(0)
    ret, ln = torch.zeros(300), torch.zeros(300, dtype=torch.int32)
    for t in range(50):
        r = torch.rand(300, generator=g)
        dn = (torch.rand(300, generator=g) < 0.1).long()
        _, _, _, info = w.step((r.to(DEV), dn.to(DEV)))
        ret += r                                       # RPO-LSTM/utils.py:22-29
        ln += 1
        assert torch.equal(info["r"].cpu(), ret) and torch.equal(info["l"].cpu(), ln)
        ret *= 1 - dn
        ln *= (1 - dn).int()
