"""CPU oracle of the estimator / controller glue of `EKFLeeLanded.pre_physics_step` -- TEST INFRASTRUCTURE.

Restates isaacgymenvs/tasks/ekf_lee_landed.py:339-530 on top of the component oracles (ahrs_ekf.EKFBank,
pv_filter.PVFilterBank, lee_control.lee_control), batched over envs:
  :345-346,366-368  accel = (v - v_prev)/dt, z += 9.8 (in place: the PV filter gets the gravity-added, un-rotated vector)
  :348-358          warm-up / reset re-seeding of Q_state and of the PV states
  :370-391          EKF update with the (possibly faulted) gyro and quaternion measurement
  :397-444          PV prediction + gated fixes; shared trigger counters => fix iff (step*N + env) % period == phase
  :458-503          carrot-waypoint logic and controller input
  :504-529          wrench = (m g thrust, torque) or the constant warm-up hover force
Sensor faults use the counter RNG streams of ouzelum_b200/csrc/companions.cu::sensor_frontend_kernel.
"""
import numpy as np

from . import philox as px
from .ahrs_ekf import EKFBank
from .lee_control import lee_control
from .pv_filter import PVFilterBank


def _fault(x, mode, prob, seed, ids, step, stream, per_env_flicker):
    f = np.float32
    if mode == 0:
        return x
    x = x.copy()
    if mode in (1, 3):
        p = f(0.1) if mode == 3 else f(prob)
        env = ids if per_env_flicker else np.array([px.GLOBAL_ENV], dtype=np.uint64)
        black = px.u01(px.draw(seed, env, step, px.P_FLICKER + (stream << 8))[0]) <= p
        x[np.broadcast_to(black, (x.shape[0],))] = 0
    if mode >= 2:
        r = px.draw(seed, ids, step, px.P_OBSNOISE + (stream << 8))
        lo = f(1.0 - float(f(prob)))
        rng = f(1.0 + float(f(prob))) - lo
        for j in range(x.shape[1]):
            x[:, j] = x[:, j] * (px.u01(r[j]) * rng + lo)
    return x


class EKFLeeGlue:
    def __init__(self, n, dt=0.01, convergence=300, pomdp_mode=0, pomdp_prob=0.0, seed=0, env_id_base=0, gravity_z=-9.81,
                 trigger=(7, 6, 3, 0), per_env_triggers=False, n_total=None):
        f = np.float32
        self.n, self.dt, self.conv = n, f(dt), convergence
        self.mode, self.prob, self.seed = pomdp_mode, pomdp_prob, seed
        self.ids = np.arange(n, dtype=np.uint64) + np.uint64(env_id_base)
        self.ekf = EKFBank(n, frequency=1 / dt)
        self.Q = np.zeros((n, 4))
        self.Q[:, 0] = 1
        self.pv = PVFilterBank(n, [1.0, 1.0, 1.0])
        self.prev_v = np.zeros((n, 3), f)
        self.waypoints = np.zeros((n, 3), f)
        self.step = 0
        self.mg = f(2.0 * -gravity_z)
        self.hover = f(-2.09 * gravity_z)
        self.trigger = trigger
        self.per_env = per_env_triggers      # True: every env counts its own steps instead of the reference's shared counters
        self.n_total = int(n_total or n)     # envs of the whole job: the shared counters advance once per env-iteration

    def pre_physics(self, root, target, reset):
        """root [N,13] f32 AFTER reset_idx, target [N,3], reset [N] bool -> (wrench [N,4], est [N,13], cmd [N,4])."""
        f = np.float32
        n, t = self.n, self.step
        warm = t < self.conv
        pos, quat, vel, angv = root[:, 0:3], root[:, 3:7], root[:, 7:10], root[:, 10:13]
        # `dv / self.dt` with dv a CUDA tensor and dt a Python float: torch-CUDA multiplies by the float32 reciprocal (BinaryDivTrueKernel.cu)
        acc = ((vel - self.prev_v) * (f(1.0) / self.dt)).astype(f)
        acc[:, 2] = acc[:, 2] + f(9.8)
        mode = 0 if warm else self.mode
        gyr = _fault(angv, mode, self.prob, self.seed, self.ids, t, 1, False)
        ang = _fault(quat, mode, self.prob, self.seed, self.ids, t, 3, True)
        acc_m = _fault(acc, mode, self.prob, self.seed, self.ids, t, 4, False)
        pos_m = _fault(pos, mode, self.prob, self.seed, self.ids, t, 5, False)
        vel_m = _fault(vel, mode, self.prob, self.seed, self.ids, t, 6, False)
        self.prev_v = vel.copy()
        sel = np.ones(n, bool) if warm else reset
        self.Q[sel] = quat[sel][:, [3, 0, 1, 2]].astype(np.float64)
        self.Q = self.ekf.update(self.Q / np.linalg.norm(self.Q, axis=1, keepdims=True), gyr.astype(np.float64),
                                 ang[:, [3, 0, 1, 2]].astype(np.float64))
        self.pv.state[reset, 0:3], self.pv.state[reset, 3:6], self.pv.state[reset, 6:9] = pos[reset], vel[reset], 0
        orient = quat if warm else self.Q.astype(f)
        self.pv.prediction_step(acc_m, orient, self.dt, flip_Qw=warm)
        k = np.full(n, t) if self.per_env else t * self.n_total + self.ids.astype(np.int64)
        pp, ph, vp, vh = self.trigger
        var = np.full(3, 0.0000001, f)
        if pp:
            self.pv.correction_step(gps_data=pos_m, gps_var=var, mask=(k % pp) == ph)
        if vp:
            self.pv.correction_step(vel_data=vel_m, vel_var=var, mask=(k % vp) == vh)      # gps_var None => R = 0
        if warm:
            self.waypoints = target.copy()
        tv = target - pos
        td = np.sqrt((tv[:, 0] * tv[:, 0] + tv[:, 1] * tv[:, 1]) + tv[:, 2] * tv[:, 2])
        wv = self.waypoints - pos
        wd = np.sqrt((wv[:, 0] * wv[:, 0] + wv[:, 1] * wv[:, 1]) + wv[:, 2] * wv[:, 2])
        if not warm:
            chk = ((wd < f(0.5)) | (wd > f(1.0))) & (wd != 0)
            rv = target - pos
            rv[:, 2] = (target[:, 2] + f(0.7)) - pos[:, 2]
            rd = np.sqrt((rv[:, 0] * rv[:, 0] + rv[:, 1] * rv[:, 1]) + rv[:, 2] * rv[:, 2])
            carrot = (rv / rd[:, None]) * f(0.75) + pos
            self.waypoints[chk] = carrot[chk]
            near = td < f(0.75)
            self.waypoints[near] = target[near]
            self.waypoints[near, 2] += f(0.09)
        cmd = np.zeros((n, 4), f)
        cmd[:, 0:3] = self.waypoints
        est = root.copy()
        if not warm:
            est[:, 0:3], est[:, 7:10] = self.pv.state[:, 0:3], self.pv.state[:, 3:6]
        if warm:
            wrench = np.zeros((n, 4), f)
            wrench[:, 0] = self.hover
        else:
            thrust, torque = lee_control(est, cmd, mode=0)
            wrench = np.concatenate([(self.mg * thrust)[:, None], torque], 1).astype(f)
        self.step += 1
        return wrench, est, cmd
