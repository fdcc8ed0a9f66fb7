"""CPU oracle of the fused x500 env step (kernel K1 `quad_step`) -- TEST INFRASTRUCTURE.

Restates, op for op in torch-CPU arithmetic (float32 = the reference's dtype; float64 available to
bound the float32 integration error), one `VecTask.step` of the reference's x500 tasks:

  isaacgymenvs/tasks/base/vec_task.py:313-359   VecTask.step  (clamp actions, pre/post hooks, timeout, clamp obs)
  isaacgymenvs/tasks/ouzelum.py:218-251         Ouzelum.pre_physics_step
  isaacgymenvs/tasks/ouzelum.py:180-190         Ouzelum.set_targets
  isaacgymenvs/tasks/ouzelum.py:192-216         Ouzelum.reset_idx
  isaacgymenvs/tasks/ouzelum.py:253-261,280-285 post_physics_step / compute_observations
  isaacgymenvs/tasks/ouzelum.py:302-332         compute_ingenuity_reward
  isaacgymenvs/utils/torch_jit_utils.py:66-71,198-208   quat_axis / quat_rotate
  isaacgymenvs/utils/POMDP.py:23-42             POMDPWrapper.observation (env-side use: tasks/landed.py:340)
  isaacgymenvs/RPO-LSTM/utils.py:20-35          RecordEpisodeStatisticsTorch.step (episode return)

`gym.simulate` (vec_task.py:335) is PhysX -- closed source, not in the reference tree.  It is
replaced by the explicit single-rigid-body integrator of SURVEY.md section 8a row P ("parity
unpinned" against PhysX; this file IS the contract the CUDA kernel is held to):
  per substep h = dt/substeps, semi-implicit Euler, wrench converted LOCAL->world once per control
  step and held, body-frame angular velocity carried across the substeps, |omega| clamped to 4*pi
  after the velocity update, closed-form quaternion update q <- normalize(q * exp(h/2 * omega_body))
  (row P's formula) with sin/cos evaluated by fixed polynomials (|h/2*omega| <= 0.032 rad: truncation
  < 1e-14, far below float32 eps) and the normalisation by one Newton step of 1/sqrt about 1.

Random draws follow oracle/philox.py (counter-based; SURVEY.md 8a row R explains why the
reference's global-generator stream cannot be reproduced).

float32 mode is written so that every +,-,*,/,sqrt is individually rounded in exactly the order the
CUDA kernel (compiled with -fmad=false) performs them; the integrator's explicit fused multiply-adds
(fmaf in the kernel) are reproduced with an exactly rounded `fma()` => the kernel is expected to match this
oracle BIT-EXACTLY, which is what makes the integer outputs (reset / timeout flags over long
roll-outs) exactly comparable.
"""
import math

import numpy as np
import torch

from . import philox as px
from . import x500

POMDP_NONE, POMDP_FLICKER, POMDP_NOISE, POMDP_FLICKER_NOISE = 0, 1, 2, 3
DR_NONE, DR_UNIFORM, DR_LOGUNIFORM, DR_GAUSSIAN = 0, 1, 2, 3
DR_SCALING, DR_ADDITIVE = 0, 1
DR_SCHED_NONE, DR_SCHED_LINEAR, DR_SCHED_CONSTANT = 0, 1, 2


def default_cfg(num_envs, **over):
    """Config mirroring cfg/task/Ouzelum.yaml + tasks/ouzelum.py constants (+ zero-default extras)."""
    dt = 0.01
    cfg = dict(
        num_envs=int(num_envs), env_id_base=0, seed=0,
        max_episode_length=2000,             # cfg/task/Ouzelum.yaml:10
        target_period=500,                   # ouzelum.py:221
        dt=dt, substeps=2, control_freq_inv=1,   # cfg/task/Ouzelum.yaml:20-21
        gravity_z=x500.GRAVITY_Z,
        clip_actions=1.0, clip_obs=5.0,      # cfg/task/Ouzelum.yaml:13-14
        thrust_rate=dt * 2000,               # ouzelum.py:237-238  (python double product, then f32)
        thrust_max=2000.0,                   # ouzelum.py:91-93
        die_dist=8.0, die_z=0.5, up_coef=5.0,    # ouzelum.py:314,326-327
        spawn_base=(0.0, 0.0, 1.0),          # ouzelum.py:150-151
        spawn_lo=(-1.5, -1.5, -0.2), spawn_range=(1.5 - (-1.5), 1.5 - (-1.5), 1.5 - (-0.2)),  # ouzelum.py:207-209
        target_scale=(10.0, 10.0, 1.0), target_off=(-5.0, -5.0, 1.0),   # ouzelum.py:183-184
        target_fixed=0,                      # 1: never resample (landing-family tasks drive the target externally)
        mass=x500.MASS, ixx=x500.IXX, iyy=x500.IYY, izz=x500.IZZ, arm=x500.ARM, com_z=x500.COM_Z,
        max_angvel=x500.MAX_ANGVEL,
        lin_drag=0.0, yaw_km=0.0,            # north-star extras, zero => reference behaviour
        fault_mode=0, fault_eff_lo=0.0, fault_eff_range=0.5,
        dr_enable=0,
        # per-parameter randomisation schema (dr_utils.py:71-132): (distribution, operation, lo|mu, hi|sigma, schedule, schedule_steps)
        # for mass, ixx, iyy, izz, arm, thrust scale, yaw_km -- default: scaling x uniform [0.8, 1.2), yaw_km not randomised
        dr=tuple((DR_UNIFORM, DR_SCALING, 0.8, 1.2, 0, 0) for _ in range(6)) + ((DR_NONE, DR_SCALING, 0.8, 1.2, 0, 0),),
        wrench_warmup_steps=0,
        pomdp_mode=POMDP_NONE, pomdp_prob=0.0, noise_sigma=0.0,
        plate_enable=0, plate_z=0.377, plate_radius=0.35, land_cutoff=0.0,
    )
    cfg.update(over)
    return cfg


def ieee_sqrt(x):
    """Correctly-rounded sqrt.  torch-CPU `sqrt` on MKL builds is NOT correctly rounded (0.6 % of float32
    inputs are 1-2 ulp off -- measured in the build container; torch-CUDA `sqrt`, what the reference
    runs on, IS correctly rounded), so the oracle takes numpy's, which is IEEE-exact."""
    return torch.from_numpy(np.sqrt(x.detach().numpy()))


def quat_rotate(q, v):
    """isaacgymenvs/utils/torch_jit_utils.py:198-208 (== isaacgym.torch_utils.quat_rotate), xyzw."""
    shape = q.shape
    q_w = q[:, -1]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(shape[0], 1, 3), v.view(shape[0], 3, 1)).squeeze(-1) * 2.0
    return a + b + c


def quat_axis(q, axis=0):
    """isaacgymenvs/utils/torch_jit_utils.py:66-71."""
    basis_vec = torch.zeros(q.shape[0], 3, dtype=q.dtype)
    basis_vec[:, axis] = 1
    return quat_rotate(q, basis_vec)


def compute_ingenuity_reward(root_positions, target_root_positions, root_quats, root_linvels, root_angvels,
                             reset_buf, progress_buf, max_episode_length, die_dist=8.0, die_z=0.5, up_coef=5.0):
    """isaacgymenvs/tasks/ouzelum.py:302-332 (die_z 0.3 for the landing family, e.g. landing.py:421-453)."""
    target_dist = ieee_sqrt(torch.square(target_root_positions - root_positions).sum(-1))
    pos_reward = 1.0 / (1.0 + target_dist * target_dist)
    ups = quat_axis(root_quats, 2)
    tiltage = torch.abs(1 - ups[..., 2])
    up_reward = up_coef / (1.0 + tiltage * tiltage)
    spinnage = torch.abs(root_angvels[..., 2])
    spinnage_reward = 1.0 / (1.0 + spinnage * spinnage)
    reward = pos_reward + pos_reward * (up_reward + spinnage_reward)
    ones = torch.ones_like(reset_buf)
    die = torch.zeros_like(reset_buf)
    die = torch.where(target_dist > die_dist, ones, die)
    die = torch.where(root_positions[..., 2] < die_z, ones, die)
    reset = torch.where(progress_buf >= max_episode_length - 1, ones, die)
    return reward, reset


def div_by_scalar(x, s):
    """`tensor / python_scalar` with torch-CUDA semantics: aten/native/cuda/BinaryDivTrueKernel.cu multiplies by the
    reciprocal computed in the tensor's dtype ("may lose one bit of precision compared to computing the division");
    torch-CPU divides.  The reference runs on CUDA, so the oracle follows the CUDA evaluation (<= 1 ulp apart)."""
    one = torch.ones((), dtype=x.dtype)
    return x * (one / torch.tensor(float(s), dtype=x.dtype))


def compute_observations(root_states, target_root_positions):
    """isaacgymenvs/tasks/ouzelum.py:280-285."""
    obs = torch.empty(root_states.shape[0], 13, dtype=root_states.dtype)
    obs[..., 0:3] = div_by_scalar(target_root_positions - root_states[:, 0:3], 3)
    obs[..., 3:7] = root_states[:, 3:7]
    obs[..., 7:10] = div_by_scalar(root_states[:, 7:10], 2)
    obs[..., 10:13] = div_by_scalar(root_states[:, 10:13], math.pi)
    return obs


def fma(a, b, c):
    """Fused multiply-add a*b + c with ONE rounding, element-wise on torch tensors (python scalars allowed).

    float32: the product of two float32 values is exact in float64 (48 <= 53 bits); the float64 sum is then rounded TO ODD
    (error-free TwoSum; if the sum is inexact and its last mantissa bit is even, step one ulp towards the exact value), and a
    round-to-odd value with >= 2 spare bits rounds to float32 exactly like the infinitely precise result (Boldo & Melquiond
    2008).  This matches CUDA's fmaf / FFMA and C's fmaf bit for bit.  float64 (error-bounding mode only): plain a*b + c."""
    ts = [t for t in (a, b, c) if isinstance(t, torch.Tensor)]
    dt = ts[0].dtype
    if dt == torch.float64:
        return a * b + c
    f = lambda t: (t.detach().numpy() if isinstance(t, torch.Tensor) else np.float32(t)).astype(np.float64)
    p = f(a) * f(b)
    cd = f(c)
    s = np.asarray(p + cd, dtype=np.float64)
    bb = s - p
    err = (p - (s - bb)) + (cd - bb)
    si = s.view(np.int64) if s.ndim else np.array(s).reshape(1).view(np.int64)
    s1 = s if s.ndim else s.reshape(1)
    inexact = np.isfinite(s1) & (np.broadcast_to(err, s1.shape) != 0)
    grow = (np.broadcast_to(err, s1.shape) > 0) == (s1 > 0)              # the exact value lies further from zero than s
    adj = np.where(inexact & ((si & 1) == 0), np.where(grow, 1, -1), 0).astype(np.int64)
    out = (si + adj).view(np.float64).astype(np.float32)
    return torch.from_numpy(out.reshape(s.shape) if s.ndim else out.reshape(()))


def quat_to_R(q):
    """xyzw unit quaternion -> rotation matrix entries r[i][j]; fixed op order shared with the kernel (quad_env.cuh)."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    x2, y2, z2 = x + x, y + y, z + z
    wx, wy, wz = x2 * w, y2 * w, z2 * w
    a, b = fma(-y2, y, 1.0), fma(-x2, x, 1.0)
    return [[fma(-z2, z, a), fma(x2, y, -wz), fma(x2, z, wy)],
            [fma(x2, y, wz), fma(-z2, z, b), fma(y2, z, -wx)],
            [fma(x2, z, -wy), fma(y2, z, wx), fma(-y2, y, b)]]


def _matvec(r, v):
    return [fma(r[i][2], v[2], fma(r[i][1], v[1], r[i][0] * v[0])) for i in range(3)]


def _matTvec(r, v):
    return [fma(r[2][i], v[2], fma(r[1][i], v[1], r[0][i] * v[0])) for i in range(3)]


def _cross(a, b):
    return [fma(a[1], b[2], -(a[2] * b[1])), fma(a[2], b[0], -(a[0] * b[2])), fma(a[0], b[1], -(a[1] * b[0]))]


class QuadStepOracle:
    """State-holding oracle: `step(actions)` == one `VecTask.step` of the Ouzelum task family."""

    def __init__(self, cfg, dtype=torch.float32, exact_trig=False):
        # float fields are rounded to float32 once, exactly as they sit in the C `ozl_cfg` struct
        r32 = lambda x: float(np.float32(x)) if isinstance(x, float) else x
        self.cfg = {k: (tuple(tuple(r32(y) for y in x) if isinstance(x, (tuple, list)) else r32(x) for x in v)
                        if isinstance(v, (tuple, list)) else r32(v)) for k, v in dict(cfg).items()}
        cfg = self.cfg
        self.dtype = dtype
        self.exact_trig = exact_trig      # float64 cross-check of the polynomial sin/cos
        n = self.n = int(cfg["num_envs"])
        z = lambda *s: torch.zeros(*s, dtype=dtype)
        self.root = z(n, 13)
        for j in range(3):
            self.root[:, j] = cfg["spawn_base"][j]
        self.root[:, 6] = 1.0
        self.thrust = z(n, 4)
        self.target = z(n, 3)
        self.target[:, 2] = 1.0           # ouzelum.py:73
        self.ep_ret = z(n)
        # per-env parameters: mass, ixx, iyy, izz, arm, thrust scale, yaw_km | fault rotor, onset, effectiveness
        self.params = z(n, 7)
        for j, k in enumerate(("mass", "ixx", "iyy", "izz", "arm")):
            self.params[:, j] = float(np.float32(cfg[k])) if dtype == torch.float32 else cfg[k]
        self.params[:, 5] = 1.0
        self.params[:, 6] = cfg["yaw_km"]
        self.fault_rotor = torch.zeros(n, dtype=torch.int64)
        self.fault_onset = torch.full((n,), 0x1FFFFFFF, dtype=torch.int64)   # never, until a reset draws one
        self.fault_eff = torch.ones(n, dtype=dtype)
        self.landed = torch.zeros(n, dtype=torch.bool)        # came within land_cutoff of the target this episode
        # VecTask.allocate_buffers (vec_task.py:254-277)
        self.obs_buf = z(n, 13)
        self.rew_buf = z(n)
        self.reset_buf = torch.ones(n, dtype=torch.int64)
        self.progress_buf = torch.zeros(n, dtype=torch.int64)
        self.timeout_buf = torch.zeros(n, dtype=torch.bool)
        self.step_count = 0
        # metrics (kernel K6): sums in float64, counts in int64
        self.msum = np.zeros(8, dtype=np.float64)
        self.mcnt = np.zeros(8, dtype=np.int64)
        self.mlanded = 0
        self.env_ids = np.arange(n, dtype=np.uint64) + np.uint64(cfg.get("env_id_base", 0))

    # ------------------------------------------------------------------ helpers
    def _c(self, v):
        """python scalar -> 0-dim tensor of the working dtype (one rounding, like a kernel constant)."""
        return torch.tensor(float(v), dtype=self.dtype)

    def _u(self, r):
        return torch.from_numpy(px.u01(r)).to(self.dtype)

    def _dr_apply(self, spec, nominal, r0, r1, step):
        """One randomised parameter -- isaacgymenvs/utils/dr_utils.py:71-132 (generate_random_samples) with the schedule driven
        by the step counter; same float32 operations as dr_apply() in ouzelum_b200/csrc/quad_env.cuh."""
        c = self._c
        dist, op, lo, hi, sched, steps = spec
        n = self.n
        if dist == DR_NONE:
            return c(nominal).expand(n).clone()
        a, b = c(lo), c(hi)
        if sched != DR_SCHED_NONE:
            inv = torch.tensor(1.0, dtype=self.dtype) / c(float(steps))
            ss = inv * c(float(min(step, steps))) if sched == DR_SCHED_LINEAR else c(0.0 if step < steps else 1.0)
            one_m = 1.0 - ss
            if op == DR_ADDITIVE:
                a, b = a * ss, b * ss
            elif dist == DR_GAUSSIAN:
                a, b = a * ss + one_m, b * ss
            else:
                a, b = a * ss + one_m, b * ss + one_m
        if dist == DR_UNIFORM:
            smp = a + (b - a) * self._u(r0)
        elif dist == DR_LOGUNIFORM:
            la, lb = math.log(float(a)), math.log(float(b))
            smp = torch.from_numpy(np.exp(la + (lb - la) * px.u01(r0).astype(np.float64))).to(self.dtype)
        else:
            u1 = ((r0 >> np.uint32(8)).astype(np.float64) + 1.0) * 5.9604644775390625e-08
            u2 = px.u01(r1).astype(np.float64)
            z = np.sqrt(-2.0 * np.log(u1)) * np.cos(6.283185307179586 * u2)
            smp = torch.from_numpy(float(a) + float(b) * z).to(self.dtype)
        return c(nominal) + smp if op == DR_ADDITIVE else c(nominal) * smp

    # ------------------------------------------------------------------ step
    def step(self, actions, target_in=None, act_mode=0, warmup=None, det_target=None):
        """act_mode 0: rotor thrust-rate actions (ouzelum.py:237-244); 1: body wrench (fz,tx,ty,tz) on the base link
        (lee_landed.py:316-330).  target_in [N,3]: externally driven target (landing.py:373-374)."""
        cfg, c, dt_ = self.cfg, self._c, self.dtype
        seed, t = cfg.get("seed", 0), self.step_count
        if warmup is None:      # estimator warm-up of the wrench-actuated classical task (ekf_lee_landed.py:339)
            warmup = act_mode == 1 and t < int(cfg.get("wrench_warmup_steps", 0))
        a = actions.to(dt_)
        if act_mode == 0:
            a = torch.clamp(a, -cfg["clip_actions"], cfg["clip_actions"])                 # vec_task.py:327
        prog, rst = self.progress_buf, self.reset_buf != 0

        # ---- pre_physics_step: target resample (ouzelum.py:221-224) and reset (ouzelum.py:226-229)
        resample = (prog % cfg["target_period"] == 0) | rst
        if cfg.get("target_fixed", 0):
            resample = torch.zeros_like(rst)
        r0, r1, r2, _ = px.draw(seed, self.env_ids, t, px.P_TARGET)
        u = [self._u(r) for r in (r0, r1, r2)]
        new_t = torch.stack([u[j] * c(cfg["target_scale"][j]) + c(cfg["target_off"][j]) for j in range(3)], -1)
        self.target = torch.where(resample[:, None], new_t, self.target)
        landed_episode = rst & self.landed                                            # landed.py:265-271
        self.landed = self.landed & ~rst
        self._landed_episode = landed_episode

        r0, r1, r2, _ = px.draw(seed, self.env_ids, t, px.P_SPAWN)
        u = [self._u(r) for r in (r0, r1, r2)]
        spawn = torch.zeros(self.n, 13, dtype=dt_)
        for j in range(3):      # ouzelum.py:206-209: initial + ((hi-lo)*rand + lo)
            spawn[:, j] = c(cfg["spawn_base"][j]) + (c(cfg["spawn_range"][j]) * u[j] + c(cfg["spawn_lo"][j]))
        spawn[:, 6] = 1.0
        self.root = torch.where(rst[:, None], spawn, self.root)
        prog = torch.where(rst, torch.zeros_like(prog), prog)                              # ouzelum.py:213-214

        # ---- north-star extras drawn at reset: rotor fault schedule, domain-randomised body parameters
        if cfg["fault_mode"]:
            r0, r1, r2, _ = px.draw(seed, self.env_ids, t, px.P_FAULT)
            rotor = torch.from_numpy((r0 & np.uint32(3)).astype(np.int64))
            onset = torch.from_numpy(px.mulhi(r1, cfg["max_episode_length"]))
            eff = c(cfg["fault_eff_lo"]) + c(cfg["fault_eff_range"]) * self._u(r2)
            self.fault_rotor = torch.where(rst, rotor, self.fault_rotor)
            self.fault_onset = torch.where(rst, onset, self.fault_onset)
            self.fault_eff = torch.where(rst, eff, self.fault_eff)
        if cfg["dr_enable"]:
            r = px.draw(seed, self.env_ids, t, px.P_DR0) + px.draw(seed, self.env_ids, t, px.P_DR1)[:3]
            r2 = px.draw(seed, self.env_ids, t, px.P_DR2) + px.draw(seed, self.env_ids, t, px.P_DR3)[:3]
            nominal = [cfg["mass"], cfg["ixx"], cfg["iyy"], cfg["izz"], cfg["arm"], 1.0, cfg["yaw_km"]]
            newp = torch.stack([self._dr_apply(cfg["dr"][j], nominal[j], r[j], r2[j], t) for j in range(7)], -1)
            self.params = torch.where(rst[:, None], newp, self.params)

        # landing detector (landed.py:288-295): pre-step position, target as left by the previous step
        cut = torch.zeros_like(rst)
        if cfg.get("land_cutoff", 0.0) > 0:
            # det_target: the point the detector measures to (LeeLanded: the controller target, lee_landed.py:305,318-322)
            d0 = (self.target if det_target is None else det_target.to(dt_)) - self.root[:, 0:3]
            cut = ieee_sqrt((d0[:, 0] * d0[:, 0] + d0[:, 1] * d0[:, 1]) + d0[:, 2] * d0[:, 2]) < c(cfg["land_cutoff"])
            if warmup:      # ekf_lee_landed.py:508-529: no flag during the estimator warm-up, hover force applied to every env
                cut = torch.zeros_like(rst)
            self.landed = self.landed | cut
        if target_in is not None:
            self.target = target_in.to(dt_).clone()
        if act_mode == 1:
            zero = torch.zeros_like(a[:, 0])
            off = cut if warmup else (rst | cut)              # forces[reset_env_ids] = 0 -- the torque tensor is NOT cleared for
            fz = torch.where(off, zero, a[:, 0])              # just-reset envs (lee_landed.py:324-325, ekf_lee_landed.py:519-520)
            tau_b = [torch.where(cut, zero, a[:, 1 + j]) for j in range(3)]
            self.thrust = torch.where(rst[:, None], torch.zeros_like(self.thrust), self.thrust)
            fault_active = torch.zeros_like(rst)
            self._simulate(None, wrench=(fz, tau_b))
            return self._post(prog, rst, fault_active)
        # ---- thrust command (ouzelum.py:237-248)
        thr = self.thrust + c(cfg["thrust_rate"]) * a
        thr = torch.max(torch.min(thr, c(cfg["thrust_max"])), c(0.0))
        force = thr.clone()
        thr = torch.where(rst[:, None], torch.zeros_like(thr), thr)
        force = torch.where(rst[:, None], torch.zeros_like(force), force)
        self.thrust = thr
        # rotor effectiveness (thrust scale from DR; single-rotor loss of effectiveness once progress >= onset)
        force = force * self.params[:, 5:6]
        force = torch.where(cut[:, None], torch.zeros_like(force), force)
        fault_active = (prog >= self.fault_onset) if cfg["fault_mode"] else torch.zeros_like(rst)
        for i in range(4):
            hit = fault_active & (self.fault_rotor == i)
            force[:, i] = torch.where(hit, force[:, i] * self.fault_eff, force[:, i])

        # ---- gym.simulate replacement
        self._simulate(force)
        return self._post(prog, rst, fault_active)

    def _post(self, prog, rst, fault_active):
        cfg, c, dt_ = self.cfg, self._c, self.dtype
        seed, t = cfg.get("seed", 0), self.step_count
        if cfg.get("plate_enable", 0):
            # landing plate (new; the reference relies on PhysX contact with the Husky's top plate)
            ddx, ddy = self.target[:, 0] - self.root[:, 0], self.target[:, 1] - self.root[:, 1]
            hit = (self.root[:, 2] < c(cfg["plate_z"])) & ((ddx * ddx + ddy * ddy) <= c(cfg["plate_radius"] ** 2))
            stopped = self.root.clone()
            stopped[:, 2] = c(cfg["plate_z"])
            stopped[:, 7:13] = 0
            self.root = torch.where(hit[:, None], stopped, self.root)
        # ---- post_physics_step (ouzelum.py:253-261)
        prog = prog + 1
        root = self.root
        obs = compute_observations(root, self.target)
        rew, reset = compute_ingenuity_reward(root[:, 0:3], self.target, root[:, 3:7], root[:, 7:10], root[:, 10:13],
                                              self.reset_buf, prog, cfg["max_episode_length"],
                                              cfg["die_dist"], cfg["die_z"], cfg["up_coef"])
        # reward is returned by the jit function in the working dtype
        self.rew_buf = rew.to(dt_)
        timeout = (prog >= cfg["max_episode_length"] - 1) & (reset != 0)                   # vec_task.py:345

        # ---- sensor-fault epilogue (utils/POMDP.py:23-42), env-side order as in tasks/landed.py:340
        mode = cfg["pomdp_mode"]
        if mode in (POMDP_NOISE, POMDP_FLICKER_NOISE) or mode == POMDP_FLICKER:
            blackout = False
            if mode in (POMDP_FLICKER, POMDP_FLICKER_NOISE):
                rf = px.draw(seed, np.array([px.GLOBAL_ENV]), t, px.P_FLICKER)[0]
                p = 0.1 if mode == POMDP_FLICKER_NOISE else cfg["pomdp_prob"]              # POMDP.py:16-18
                blackout = bool(px.u01(rf)[0] <= np.float32(p))                            # POMDP.py:25,33
            if blackout:
                obs = torch.zeros_like(obs)
            if mode in (POMDP_NOISE, POMDP_FLICKER_NOISE):
                us = []
                for k in range(4):
                    us += list(px.draw(seed, self.env_ids, t, px.P_OBSNOISE + k))
                sig = cfg["noise_sigma"]
                lo = c(1.0 - sig)
                rng = c(1.0 + sig) - lo            # uniform_(from, to): (to - from) in the tensor dtype
                noise = torch.stack([self._u(us[j]) * rng + lo for j in range(13)], -1)   # uniform_(lo, hi)
                obs = obs * noise
        self.obs_buf = torch.clamp(obs, -cfg["clip_obs"], cfg["clip_obs"])                # vec_task.py:353

        # ---- episode statistics (RPO-LSTM/utils.py:20-35) + metrics (K6)
        ep_ret = self.ep_ret + self.rew_buf
        done = reset != 0
        d = (self.target - root[:, 0:3])
        dist = ieee_sqrt(torch.square(d).sum(-1))
        self.msum[0] += float(self.rew_buf.double().sum())
        self.msum[1] += float(ep_ret[done].double().sum())
        self.mcnt[0] += self.n
        self.mcnt[1] += int(done.sum())
        self.mcnt[2] += int(prog[done].sum())
        self.mcnt[3] += int(timeout.sum())
        self.mcnt[4] += int((dist > cfg["die_dist"]).sum())
        self.mcnt[5] += int((root[:, 2] < cfg["die_z"]).sum())
        self.mcnt[6] += int(fault_active.sum())
        self.mcnt[7] += int(rst.sum())
        self.mlanded += int(self._landed_episode.sum())
        self.returned_ep_ret = ep_ret.clone()
        self.ep_ret = torch.where(done, torch.zeros_like(ep_ret), ep_ret)

        self.progress_buf, self.reset_buf, self.timeout_buf = prog, reset, timeout
        self.step_count += 1
        return self.obs_buf, self.rew_buf, self.reset_buf, self.timeout_buf

    # ------------------------------------------------------------------ rigid body
    def _simulate(self, force, wrench=None, body_force=None):
        """Integrator "row P" v2 -- same operations, same order as `simulate()` in ouzelum_b200/csrc/quad_env.cuh."""
        cfg, c = self.cfg, self._c
        root, P = self.root, self.params
        p = [root[:, j] for j in range(0, 3)]
        q = root[:, 3:7].clone()
        v = [root[:, j] for j in range(7, 10)]
        w = [root[:, j] for j in range(10, 13)]
        inv_m = 1.0 / P[:, 0]
        inertia = [P[:, 1], P[:, 2], P[:, 3]]
        h = c(cfg["dt"] / cfg["substeps"])
        hh = c(0.5 * (cfg["dt"] / cfg["substeps"]))
        hh2 = hh * hh
        hi = [h * (1.0 / P[:, 1]), h * (1.0 / P[:, 2]), h * (1.0 / P[:, 3])]
        arm, cz = P[:, 4], c(cfg["com_z"])
        if wrench is not None:
            fz, tau_b = wrench
        else:
            f0, f1, f2, f3 = force[:, 0], force[:, 1], force[:, 2], force[:, 3]
            fz = ((f0 + f1) + f2) + f3
            tau_b = [arm * (((f1 - f0) + f2) - f3),
                     arm * (((f1 - f0) - f2) + f3),
                     P[:, 6] * (((f2 - f0) - f1) + f3)]
        R = quat_to_R(q)
        b3 = [R[0][2], R[1][2], R[2][2]]
        fw = _matvec(R, body_force) if body_force is not None else [b3[j] * fz for j in range(3)]
        tau_w = _matvec(R, tau_b)
        g = [c(0.0), c(0.0), c(cfg["gravity_z"])]
        kdm = c(cfg["lin_drag"]) * inv_m
        aw = [fma(fw[j], inv_m, g[j]) for j in range(3)]
        # root (base-link origin) -> composite COM
        rc = [cz * b3[j] for j in range(3)]
        x = [p[j] + rc[j] for j in range(3)]
        wxr = _cross(w, rc)
        v = [v[j] + wxr[j] for j in range(3)]
        wb = _matTvec(R, w)
        tb = list(tau_b)
        wmax, wmax2 = c(cfg["max_angvel"]), c(cfg["max_angvel"] ** 2)
        nsub = int(cfg["substeps"]) * int(cfg["control_freq_inv"])
        one = torch.ones_like(inv_m)
        for s_ in range(nsub):
            # linear velocity (world), position
            for j in range(3):
                v[j] = fma(h, fma(-kdm, v[j], aw[j]), v[j])
                x[j] = fma(h, v[j], x[j])
            # angular velocity: Euler's equations in the body frame
            iw = [inertia[j] * wb[j] for j in range(3)]
            gy = _cross(wb, iw)
            wb = [fma(hi[j], tb[j] - gy[j], wb[j]) for j in range(3)]
            n2 = fma(wb[2], wb[2], fma(wb[1], wb[1], wb[0] * wb[0]))
            over = n2 > wmax2
            scale = torch.where(over, wmax / ieee_sqrt(n2), one)
            wb = [torch.where(over, wb[j] * scale, wb[j]) for j in range(3)]
            n2 = torch.where(over, fma(wb[2], wb[2], fma(wb[1], wb[1], wb[0] * wb[0])), n2)
            # attitude: q <- normalize(q (x) (k w_b, cos))
            th2 = hh2 * n2                                        # (h/2 |w|)^2
            if self.exact_trig:
                th = ieee_sqrt(th2)
                sinc = torch.where(th > 0, torch.sin(th) / torch.where(th > 0, th, torch.ones_like(th)), torch.ones_like(th))
                cs = torch.cos(th)
            else:
                sinc = fma(th2, fma(th2, c(1.0 / 120.0), c(-1.0 / 6.0)), 1.0)
                cs = fma(th2, fma(th2, fma(th2, c(-1.0 / 720.0), c(1.0 / 24.0)), c(-0.5)), 1.0)
            k = hh * sinc
            dx, dy, dz = k * wb[0], k * wb[1], k * wb[2]
            qx, qy, qz, qw = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
            nx = fma(qw, dx, fma(cs, qx, fma(qy, dz, -(qz * dy))))
            ny = fma(qw, dy, fma(cs, qy, fma(qz, dx, -(qx * dz))))
            nz = fma(qw, dz, fma(cs, qz, fma(qx, dy, -(qy * dx))))
            nw = fma(qw, cs, -fma(qx, dx, fma(qy, dy, qz * dz)))
            s2 = fma(nx, nx, fma(ny, ny, fma(nz, nz, nw * nw)))
            if self.exact_trig:
                inv = 1.0 / ieee_sqrt(s2)
            else:
                inv = fma(c(-0.5), s2, 1.5)                       # one Newton step of 1/sqrt about s2 = 1
            q = torch.stack([nx * inv, ny * inv, nz * inv, nw * inv], -1)
            R = quat_to_R(q)
            if s_ + 1 < nsub:
                tb = _matTvec(R, tau_w)                           # held world-frame torque seen from the new attitude
        # body rates -> world; composite COM -> root (base-link origin): p = x - R c ; v_root = v_com - w x (R c)
        w = _matvec(R, wb)
        b3 = [R[0][2], R[1][2], R[2][2]]
        rc = [cz * b3[j] for j in range(3)]
        wxr = _cross(w, rc)
        out = torch.empty_like(root)
        for j in range(3):
            out[:, j] = x[j] - rc[j]
            out[:, 7 + j] = v[j] - wxr[j]
            out[:, 10 + j] = w[j]
        out[:, 3:7] = q
        self.root = out

    # ------------------------------------------------------------------ state access (parity tests)
    def post_reset_root(self):
        """Root state as `reset_idx` leaves it for the envs flagged in reset_buf (spawn draws of the CURRENT step), without
        stepping: what the estimator / controller code of the classical tasks sees (reset_idx runs first in
        pre_physics_step, ekf_lee_landed.py:312-314)."""
        cfg, c = self.cfg, self._c
        rst = self.reset_buf != 0
        r0, r1, r2, _ = px.draw(cfg.get("seed", 0), self.env_ids, self.step_count, px.P_SPAWN)
        u = [self._u(r) for r in (r0, r1, r2)]
        spawn = torch.zeros(self.n, 13, dtype=self.dtype)
        for j in range(3):
            spawn[:, j] = c(cfg["spawn_base"][j]) + (c(cfg["spawn_range"][j]) * u[j] + c(cfg["spawn_lo"][j]))
        spawn[:, 6] = 1.0
        return torch.where(rst[:, None], spawn, self.root)

    def load(self, root, thrust, target, ep_ret, params8, fault2, reset_buf, progress_buf, step_count):
        """Adopt the device-side state of an env handle (ozl_get_state / ozl_get_params layouts) so that ONE step can be compared
        from identical state.  params8 = mass, ixx, iyy, izz, arm, thrust scale, fault effectiveness, yaw_km; fault2 = rotor,
        onset | landed << 31."""
        t = lambda x, dt=None: torch.as_tensor(np.asarray(x)).to(dt or self.dtype).clone()
        self.root, self.thrust, self.target, self.ep_ret = t(root), t(thrust), t(target), t(ep_ret)
        p = t(params8)
        self.params = torch.cat([p[:, 0:6], p[:, 7:8]], 1)
        self.fault_eff = p[:, 6].clone()
        f = torch.as_tensor(np.asarray(fault2)).to(torch.int64)
        self.fault_rotor = f[:, 0].clone()
        self.fault_onset = f[:, 1] & 0x1FFFFFFF
        self.landed = f[:, 1] < 0
        self.reset_buf, self.progress_buf = t(reset_buf, torch.int64), t(progress_buf, torch.int64)
        self.step_count = int(step_count)

    def get_state(self):
        return dict(root=self.root.clone(), thrust=self.thrust.clone(), target=self.target.clone(),
                    ep_ret=self.ep_ret.clone(), params=self.params.clone(),
                    fault_rotor=self.fault_rotor.clone(), fault_onset=self.fault_onset.clone(),
                    fault_eff=self.fault_eff.clone())
