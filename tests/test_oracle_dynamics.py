"""Properties of the rigid-body integrator the kernel is held to, CPU only: float32 stays within the
stated tolerance of float64 over a fixed 200-step horizon, the polynomial sin/cos matches the exact trig, analytic cases."""
import numpy as np
import torch

from oracle.quad_step import QuadStepOracle, default_cfg
from oracle import x500


def _bounded_actions(n, steps, seed):
    g = torch.Generator().manual_seed(seed)
    return [(torch.rand(n, 4, generator=g) * 2 - 1) * 0.02 for _ in range(steps)]


def test_200_step_horizon_float32_vs_float64_position_and_attitude():
    """Tolerance stated in BASELINE north_star / SURVEY 8c: position <= 1e-3 m, attitude <= 1e-3 rad over 200 steps of
    bounded random actions from a hover-ish start."""
    n = 256
    cfg = default_cfg(n, max_episode_length=100000, die_dist=1e9, die_z=-1e9)
    o32, o64 = QuadStepOracle(cfg), QuadStepOracle(cfg, dtype=torch.float64)
    z = torch.zeros(n, 4)
    o32.step(z), o64.step(z.double())
    hover = x500.MASS * 9.81 / 4
    for o in (o32, o64):
        o.thrust[:] = hover
        o.root[:, 7:13] = 0
    for a in _bounded_actions(n, 200, 1):
        o32.step(a)
        o64.step(a.double())
    assert torch.equal(o32.progress_buf, o64.progress_buf) and int(o32.progress_buf[0]) == 201
    dp = (o32.root[:, 0:3].double() - o64.root[:, 0:3]).norm(dim=1).max().item()
    q32, q64 = o32.root[:, 3:7].double(), o64.root[:, 3:7]
    ang = 2 * torch.acos(torch.clamp((q32 * q64).sum(1).abs(), max=1.0)).max().item()
    assert dp < 1e-3, dp
    assert ang < 1e-3, ang
    assert (o64.root[:, 0:3] - torch.tensor(cfg["spawn_base"], dtype=torch.float64)).norm(dim=1).max() > 0.05   # it did move


def test_polynomial_trig_equals_exact_trig_in_float64():
    n = 128
    cfg = default_cfg(n)
    a, b = QuadStepOracle(cfg, dtype=torch.float64), QuadStepOracle(cfg, dtype=torch.float64, exact_trig=True)
    g = torch.Generator().manual_seed(2)
    for _ in range(60):
        act = (torch.rand(n, 4, generator=g, dtype=torch.float64) * 2 - 1)
        a.step(act), b.step(act)
    same = a.progress_buf == b.progress_buf
    assert same.all()
    assert (a.root - b.root).abs().max() < 1e-9


def test_free_fall_and_pure_yaw_torque_analytic():
    cfg = default_cfg(4, die_dist=1e9, die_z=-1e9, yaw_km=0.1)
    o = QuadStepOracle(cfg, dtype=torch.float64)
    o.step(torch.zeros(4, 4, dtype=torch.float64))                  # reset step: thrust zeroed -> free fall for one step
    v1 = o.root[:, 9].clone()
    assert torch.allclose(v1, torch.full((4,), -9.81 * 0.01, dtype=torch.float64), atol=1e-6)      # f32-rounded g and dt
    z1 = o.root[:, 2].clone()
    o.step(torch.zeros(4, 4, dtype=torch.float64))
    # semi-implicit Euler, 2 substeps of 5 ms: dz = h*(v + h g) + h*(v + 2 h g)
    h, g = 0.005, -9.81
    assert torch.allclose(o.root[:, 2] - z1, 2 * h * v1 + 3 * h * h * g, atol=1e-7)
    # rotors 2,3 only -> pure reaction torque about body z (ccw pair off): yaw rate grows, no roll/pitch
    o.thrust[:] = torch.tensor([0.0, 0.0, 5.0, 5.0], dtype=torch.float64)
    o.root[:, 10:13] = 0
    o.step(torch.zeros(4, 4, dtype=torch.float64))
    w = o.root[0, 10:13]
    assert abs(w[2] - 0.1 * 10.0 / float(np.float32(x500.IZZ)) * 0.01) < 1e-6
    # rotors 2 and 3 sit at (+x,+y) and (-x,-y): their roll/pitch torques cancel
    assert abs(w[0]) < 1e-9 and abs(w[1]) < 1e-9


def test_reset_ordering_quirk_and_target_resample():
    """an env flagged at step t is re-initialised at the START of step t+1, still gets one physics step with
    zero thrust, and reports progress == 1; the target is re-drawn whenever progress % 500 == 0 (ouzelum.py:221-224)."""
    cfg = default_cfg(8, max_episode_length=20)
    o = QuadStepOracle(cfg)
    a = torch.zeros(8, 4)
    o.step(a)
    assert (o.progress_buf == 1).all() and (o.thrust == 0).all()
    t0 = o.target.clone()
    for _ in range(18):
        o.step(a + 0.25)
    assert (o.reset_buf == 1).all() and (o.timeout_buf).all() and (o.progress_buf == 19).all()
    o.step(a + 1.0)
    assert (o.progress_buf == 1).all() and (o.thrust == 0).all() and not torch.equal(o.target, t0)
    cfg = default_cfg(4, target_period=5, die_dist=1e9, die_z=-1e9)
    o = QuadStepOracle(cfg)
    tg = []
    for _ in range(12):
        o.step(torch.zeros(4, 4))
        tg.append(o.target.clone())
    changes = [k for k in range(1, 12) if not torch.equal(tg[k], tg[k - 1])]
    assert changes == [5, 10]            # progress 5 and 10 at the start of steps 6 and 11 (0-based 5, 10)


def test_torque_free_tumbling_conserves_angular_momentum_and_energy():
    """Row P is a restatement without a source to follow (PhysX), so it is also held to the physics: with no thrust and no
    gravity a tumbling body keeps its world-frame angular momentum L = R I omega_b and its rotational energy up to the drift
    of a first-order scheme (semi-implicit Euler, h = 5 ms), the drift shrinks with the substep (first-order convergence), the
    quaternion stays normalised, and the centre of mass moves in a straight line."""
    n = 64
    inertia = torch.tensor([float(np.float32(x500.IXX)), float(np.float32(x500.IYY)), float(np.float32(x500.IZZ))], dtype=torch.float64)
    com_z = float(np.float32(x500.COM_Z))

    def rot(qq):
        x, y, zz, w = qq[:, 0], qq[:, 1], qq[:, 2], qq[:, 3]
        return torch.stack([torch.stack([1 - 2 * (y * y + zz * zz), 2 * (x * y - w * zz), 2 * (x * zz + w * y)], 1),
                            torch.stack([2 * (x * y + w * zz), 1 - 2 * (x * x + zz * zz), 2 * (y * zz - w * x)], 1),
                            torch.stack([2 * (x * zz - w * y), 2 * (y * zz + w * x), 1 - 2 * (x * x + y * y)], 1)], 1)

    def run(substeps):
        cfg = default_cfg(n, die_dist=1e9, die_z=-1e9, gravity_z=0.0, max_angvel=1e6, substeps=substeps)
        o = QuadStepOracle(cfg, dtype=torch.float64)
        z = torch.zeros(n, 4, dtype=torch.float64)
        o.step(z)                                                             # applies the initial reset
        g = torch.Generator().manual_seed(4)
        q = torch.randn(n, 4, generator=g, dtype=torch.float64)
        o.root[:, 3:7] = q / q.norm(dim=1, keepdim=True)
        o.root[:, 7:10] = torch.randn(n, 3, generator=g, dtype=torch.float64)
        o.root[:, 10:13] = torch.randn(n, 3, generator=g, dtype=torch.float64) * 3.0   # |omega| ~ 5 rad/s, Izz != Ixx

        def invariants():
            R = rot(o.root[:, 3:7])
            wb = torch.einsum("nji,nj->ni", R, o.root[:, 10:13])              # body rates = R^T omega_world
            L = torch.einsum("nij,nj->ni", R, inertia * wb)
            E = 0.5 * (inertia * wb * wb).sum(1)
            com = o.root[:, 0:3] + com_z * R[:, :, 2]                         # root -> composite centre of mass
            return L, E, com
        L0, E0, c0 = invariants()
        v_com0 = None
        for t in range(200):
            o.step(z)
            if t == 0:
                v_com0 = (invariants()[2] - c0) / 0.01
        L1, E1, c200 = invariants()
        assert (o.root[:, 3:7].norm(dim=1) - 1).abs().max() < 1e-12
        assert ((c200 - c0) - v_com0 * 2.0).norm(dim=1).max() < 1e-9, "the centre of mass moves uniformly"
        return ((L1 - L0).norm(dim=1) / L0.norm(dim=1)).max(), ((E1 - E0).abs() / E0).max()
    dl2, de2 = run(2)
    dl8, de8 = run(8)
    assert dl2 < 5e-2 and de2 < 5e-2, (dl2, de2)              # 2 s of fast tumbling at the task's substep
    assert dl8 < 0.4 * dl2 and de8 < 0.4 * de2, (dl2, dl8, de2, de8)      # 4x smaller substep: ~4x smaller drift (first order)
