"""CPU oracle of the stock Quadcopter hover task's observation / reward / reset (kernel K1q) -- TEST INFRASTRUCTURE.

Restates isaacgymenvs/tasks/quadcopter.py:359-370 (compute_observations) and :386-418 (compute_quadcopter_reward).
The articulated 9-body PhysX dynamics of that task are out of parity scope.
"""
import math

import torch

from .quad_step import div_by_scalar, ieee_sqrt, quat_axis


def compute_quadcopter_reward(root_positions, root_quats, root_linvels, root_angvels, reset_buf, progress_buf, max_episode_length):
    """quadcopter.py:386-418: target (0,0,1), up_reward 1/(1+tilt^2), die dist>3 or z<0.3."""
    target_dist = ieee_sqrt(root_positions[..., 0] * root_positions[..., 0] +
                            root_positions[..., 1] * root_positions[..., 1] +
                            (1 - root_positions[..., 2]) * (1 - root_positions[..., 2]))
    pos_reward = 1.0 / (1.0 + target_dist * target_dist)
    ups = quat_axis(root_quats, 2)
    tiltage = torch.abs(1 - ups[..., 2])
    up_reward = 1.0 / (1.0 + tiltage * tiltage)
    spinnage = torch.abs(root_angvels[..., 2])
    spinnage_reward = 1.0 / (1.0 + spinnage * spinnage)
    reward = pos_reward + pos_reward * (up_reward + spinnage_reward)
    ones = torch.ones_like(reset_buf)
    die = torch.zeros_like(reset_buf)
    die = torch.where(target_dist > 3.0, ones, die)
    die = torch.where(root_positions[..., 2] < 0.3, ones, die)
    reset = torch.where(progress_buf >= max_episode_length - 1, ones, die)
    return reward, reset


def compute_observations(root_states, dof_positions):
    """quadcopter.py:359-370: 21 observations."""
    n = root_states.shape[0]
    obs = torch.empty(n, 21, dtype=root_states.dtype)
    obs[..., 0] = div_by_scalar(0.0 - root_states[..., 0], 3)
    obs[..., 1] = div_by_scalar(0.0 - root_states[..., 1], 3)
    obs[..., 2] = div_by_scalar(1.0 - root_states[..., 2], 3)
    obs[..., 3:7] = root_states[..., 3:7]
    obs[..., 7:10] = div_by_scalar(root_states[..., 7:10], 2)
    obs[..., 10:13] = div_by_scalar(root_states[..., 10:13], math.pi)
    obs[..., 13:21] = dof_positions
    return obs


# ---------------------------------------------------------------------------------------------------- full step oracle
import numpy as np  # noqa: E402

from . import philox as px  # noqa: E402
from .quad_step import QuadStepOracle, _cross, default_cfg  # noqa: E402


def vehicle_constants():
    """Composite rigid body of the procedurally built vehicle (quadcopter.py:121-202), zero tilt:
    chassis cylinder r 0.1, h 0.03, density 50; 4 arm spheres r 0.01, density 200 at radius 0.1025;
    4 rotor cylinders r 0.04, h 0.01, density 1000 at radius 0.145; arms at 45/135/225/315 degrees."""
    pi = math.pi
    m_c = pi * 0.1 ** 2 * 0.03 * 50
    m_a = 4.0 / 3.0 * pi * 0.01 ** 3 * 200
    m_r = pi * 0.04 ** 2 * 0.01 * 1000
    ra, rr = 0.1 + 0.25 * 0.01, (0.1 + 0.25 * 0.01) + (0.04 + 0.25 * 0.01)
    ixx = m_c * (3 * 0.1 ** 2 + 0.03 ** 2) / 12 + 4 * (0.4 * m_a * 0.01 ** 2) + 2 * m_a * ra ** 2 \
        + 4 * (m_r * (3 * 0.04 ** 2 + 0.01 ** 2) / 12) + 2 * m_r * rr ** 2
    izz = 0.5 * m_c * 0.1 ** 2 + 4 * (0.4 * m_a * 0.01 ** 2) + 4 * m_a * ra ** 2 + 4 * (0.5 * m_r * 0.04 ** 2) + 4 * m_r * rr ** 2
    return dict(mass=m_c + 4 * m_a + 4 * m_r, ixx=ixx, iyy=ixx, izz=izz)


def _sincos_small(x):
    """Fixed polynomials shared with ouzelum_b200/csrc/quadcopter.cu (|x| <= 0.6)."""
    f = np.float32
    c = lambda v: torch.tensor(float(f(v)), dtype=x.dtype)
    x2 = x * x
    s = x * (1.0 + x2 * (c(-1.6666667e-1) + x2 * (c(8.3333333e-3) + x2 * (c(-1.9841270e-4) + x2 * c(2.7557319e-6)))))
    co = 1.0 + x2 * (c(-0.5) + x2 * (c(4.1666667e-2) + x2 * (c(-1.3888889e-3) + x2 * (c(2.4801587e-5) + x2 * c(-2.7557319e-7)))))
    return s, co


class QuadcopterOracle:
    """One `VecTask.step` of tasks/quadcopter.py with the single-rigid-body / kinematic-tilt dynamics of K1q."""

    def __init__(self, n, seed=0, env_id_base=0, max_episode_length=500, dt=0.01, substeps=2, clip_actions=1.0, clip_obs=5.0,
                 dtype=torch.float32):
        vc = vehicle_constants()
        cfg = default_cfg(n, seed=seed, env_id_base=env_id_base, max_episode_length=max_episode_length, dt=dt,
                          substeps=substeps, com_z=0.0, **vc)
        self.body = QuadStepOracle(cfg, dtype=dtype)          # reuse the integrator + float32 config rounding
        self.cfg, self.n, self.dtype = self.body.cfg, n, dtype
        self.clip_actions, self.clip_obs = clip_actions, clip_obs
        z = lambda *s: torch.zeros(*s, dtype=dtype)
        self.root = z(n, 13)
        self.root[:, 2] = 1.0
        self.root[:, 6] = 1.0
        self.dof_pos, self.dof_tgt, self.thrust = z(n, 8), z(n, 8), z(n, 4)
        self.obs_buf, self.rew_buf = z(n, 21), z(n)
        self.reset_buf = torch.ones(n, dtype=torch.int64)
        self.progress_buf = torch.zeros(n, dtype=torch.int64)
        self.timeout_buf = torch.zeros(n, dtype=torch.bool)
        self.step_count = 0
        self.ids = np.arange(n, dtype=np.uint64) + np.uint64(env_id_base)

    def step(self, actions):
        b, cfg, dt_ = self.body, self.cfg, self.dtype
        c, u = b._c, b._u
        t, seed = self.step_count, cfg["seed"]
        a = torch.clamp(actions.to(dt_), -self.clip_actions, self.clip_actions)
        rst = self.reset_buf != 0
        prog = torch.where(rst, torch.zeros_like(self.progress_buf), self.progress_buf)
        # reset_idx (quadcopter.py:280-299)
        r0, r1, r2, _ = px.draw(seed, self.ids, t, px.P_SPAWN)
        spawn = torch.zeros(self.n, 13, dtype=dt_)
        for j, r in enumerate((r0, r1, r2)):
            spawn[:, j] = c(cfg["spawn_base"][j]) + (c(cfg["spawn_range"][j]) * u(r) + c(cfg["spawn_lo"][j]))
        spawn[:, 6] = 1.0
        self.root = torch.where(rst[:, None], spawn, self.root)
        rr = list(px.draw(seed, self.ids, t, px.P_QDOF0)) + list(px.draw(seed, self.ids, t, px.P_QDOF1))
        newdp = torch.stack([c(0.4) * u(r) + c(-0.2) for r in rr], -1)
        self.dof_pos = torch.where(rst[:, None], newdp, self.dof_pos)
        # pre_physics_step (quadcopter.py:301-330)
        dof_rate = c(cfg["dt"] * 8.0 * math.pi)
        lim = c(30.0 * math.pi / 180.0)
        tg = self.dof_tgt + dof_rate * a[:, 0:8]
        tg = torch.max(torch.min(tg, lim), -lim)
        self.dof_tgt = torch.where(rst[:, None], self.dof_pos, tg)
        th = self.thrust + c(cfg["dt"] * 200.0) * a[:, 8:12]
        th = torch.max(torch.min(th, c(2.0)), c(0.0))
        force = torch.where(rst[:, None], torch.zeros_like(th), th)
        self.thrust = force.clone()
        # wrench of the four tilting rotors
        arm_r, rotor_off = c(0.1 + 0.25 * 0.01), c(0.04 + 0.25 * 0.01)
        fb = [torch.zeros(self.n, dtype=dt_) for _ in range(3)]
        tau = [torch.zeros(self.n, dtype=dt_) for _ in range(3)]
        for k in range(4):
            sp, cp = _sincos_small(self.dof_tgt[:, 2 * k])
            sr, cr = _sincos_small(self.dof_tgt[:, 2 * k + 1])
            dl = [sp * cr, -sr, cp * cr]
            pl = [arm_r + rotor_off * cp, torch.zeros_like(sp), -rotor_off * sp]
            ang = (0.25 + 0.5 * k) * math.pi
            ca, sa = c(math.cos(ang)), c(math.sin(ang))
            d = [ca * dl[0] - sa * dl[1], sa * dl[0] + ca * dl[1], dl[2]]
            p = [ca * pl[0] - sa * pl[1], sa * pl[0] + ca * pl[1], pl[2]]
            f = [force[:, k] * d[j] for j in range(3)]
            t3 = _cross(p, f)                                   # same fused form as cross3() in quad_env.cuh
            fb = [fb[j] + f[j] for j in range(3)]
            tau = [tau[j] + t3[j] for j in range(3)]
        b.root = self.root
        b._simulate(None, wrench=(None, tau), body_force=fb)
        self.root = b.root
        self.dof_pos = self.dof_tgt.clone()
        # post_physics_step
        prog = prog + 1
        obs = compute_observations(self.root, self.dof_pos)
        rew, reset = compute_quadcopter_reward(self.root[:, 0:3], self.root[:, 3:7], self.root[:, 7:10], self.root[:, 10:13],
                                               self.reset_buf, prog, cfg["max_episode_length"])
        self.obs_buf = torch.clamp(obs, -self.clip_obs, self.clip_obs)
        self.rew_buf = rew.to(dt_)
        self.timeout_buf = (prog >= cfg["max_episode_length"] - 1) & (reset != 0)
        self.progress_buf, self.reset_buf = prog, reset
        self.step_count += 1
        return self.obs_buf, self.rew_buf, self.reset_buf, self.timeout_buf
