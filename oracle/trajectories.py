"""CPU oracle of the waypoint tables and the Husky waypoint follower (kernel K5) -- TEST INFRASTRUCTURE.

Restates
  isaacgymenvs/utils/trajectories.py:5-17   lemniscate      :19-29 circle      :31-60 square
  isaacgymenvs/utils/controllers.py:5-13    map_to_pi       :15-43 differential_drive
  isaacgymenvs/tasks/landing.py:208-213     per-env trajectory id / scale / direction / index
  isaacgymenvs/tasks/landing.py:224-244     reset_completed_trajectories
  isaacgymenvs/tasks/landing.py:319-364     set_husky_actions (waypoint state machine)
  isaacgym.torch_utils.get_euler_xyz        third-party (Isaac Gym Preview, not in the reference tree): published formula
Tables + differential_drive are pinned against the reference functions run on CPU (tests/golden/trajectories.npz,
diffdrive.npz).  The Husky itself is PhysX in the reference; its replacement (kinematic unicycle driven by the
wheel speeds) is new and "parity unpinned" (SURVEY.md 8a row G3).
"""
import math

import numpy as np

from . import philox as px

WHEEL_BASE, WHEEL_RADIUS, MAX_SPEED = 0.54, 0.165, 15     # controllers.py:18-20
NUM_WAYPOINTS = 100                                        # landing.py:108


def lemniscate(a=math.sqrt(2), num_points=200):
    """trajectories.py:5-17 (torch.linspace float32 arithmetic)."""
    f = np.float32
    n = num_points
    # torch.linspace(start, end, n) in float32: start + i*step for the first half, end - (n-1-i)*step for the second
    start, end = f(-math.pi / 2), f(3 * math.pi / 2)
    step = (end - start) / f(n - 1)
    i = np.arange(n)
    theta = np.where(i < n // 2, start + step * i.astype(f), end - step * (n - 1 - i).astype(f)).astype(f)
    s, c = np.sin(theta), np.cos(theta)
    x = f(a) * c / (s ** 2 + f(1))
    y = f(a) * c * s / (s ** 2 + f(1))
    return np.stack([x, y], -1).astype(f)


def circle(r=math.sqrt(2), num_points=200):
    """trajectories.py:19-29: python float64 math, then torch.tensor -> float32."""
    step = 360 / num_points
    pts = [(r * math.cos(math.radians(i * step)), r * math.sin(math.radians(i * step))) for i in range(num_points)]
    return np.asarray(pts, dtype=np.float32)


def square(side_length=5, num_points=8):
    """trajectories.py:31-60."""
    if num_points < 4:
        raise ValueError("A square needs at least 4 waypoints.")
    per = num_points // 4
    inc = side_length / (per - 1)
    w = [(i * inc, 0) for i in range(per)]
    w += [(side_length, i * inc) for i in range(1, per)]
    w += [(side_length - i * inc, side_length) for i in range(1, per)]
    w += [(0, side_length - i * inc) for i in range(1, per - 1)]
    return -(np.asarray(w, dtype=np.float32) - np.float32(side_length / 2))


def landing_tables():
    """landing.py:108-112."""
    return lemniscate(4, NUM_WAYPOINTS), circle(2, NUM_WAYPOINTS), square(4, 8)


def map_to_pi(angle):
    """controllers.py:5-13."""
    f = angle.dtype.type
    angle = np.where(angle > f(np.pi), angle - f(2 * np.pi), angle)
    angle = np.where(angle <= f(-np.pi), angle + f(2 * np.pi), angle)
    if np.any((angle > f(np.pi)) | (angle < f(-np.pi))):
        raise RuntimeError(f"Angle out of bounds: {angle}")
    return angle


def differential_drive(current_pos, target_pos, current_heading, p_gain=(0.5, 10), ang_thresh=0.005):
    """controllers.py:15-43 -> wheel speeds [N,4] = (right, left, right, left)."""
    f = np.float32
    cp, tp, hd = (np.asarray(a, dtype=f) for a in (current_pos, target_pos, current_heading))
    dx, dy = tp[:, 0] - cp[:, 0], tp[:, 1] - cp[:, 1]
    dth = map_to_pi(np.arctan2(dy, dx) - map_to_pi(hd))
    dth = np.where((dth < f(ang_thresh)) & (dth > f(-ang_thresh)), f(0), dth)
    lin = np.sqrt(dx ** 2 + dy ** 2) * f(p_gain[0])
    ang = dth * f(p_gain[1])
    left = (f(2) * lin + ang * f(WHEEL_BASE)) / f(2 * WHEEL_RADIUS)
    right = (f(2) * lin - ang * f(WHEEL_BASE)) / f(2 * WHEEL_RADIUS)
    mx = np.maximum(np.abs(left), np.abs(right))
    with np.errstate(divide="ignore", invalid="ignore"):
        sc = f(MAX_SPEED) / mx
    over = mx > f(MAX_SPEED)
    left = np.where(over, left * sc, left)
    right = np.where(over, right * sc, right)
    return np.stack([right, left, right, left], -1).astype(f)


def get_euler_xyz(q):
    """isaacgym.torch_utils.get_euler_xyz (xyzw in, angles mod 2*pi out) -- third party, published formula."""
    f = q.dtype.type
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    roll = np.arctan2(f(2.0) * (w * x + y * z), w * w - x * x - y * y + z * z)
    sinp = f(2.0) * (w * y - z * x)
    pitch = np.where(np.abs(sinp) >= 1, np.copysign(f(np.pi / 2.0), sinp), np.arcsin(np.clip(sinp, -1, 1)))
    yaw = np.arctan2(f(2.0) * (w * z + x * y), w * w + x * x - y * y - z * z)
    two_pi = f(2 * np.pi)
    return np.remainder(roll, two_pi), np.remainder(pitch, two_pi), np.remainder(yaw, two_pi)


class HuskyFollower:
    """Waypoint state machine of landing.py:208-244,319-364 + a kinematic differential-drive vehicle (CPU twin of
    ouzelum_b200/csrc/targets.cu).

    Reference randomness (torch.randint / torch.rand on the global generator) is replaced by the counter RNG:
    a (re)draw at global step t uses philox.draw(seed, env, t, P_HUSKY): traj = mulhi(r0, 3), scale = 0.8 + 0.4*u(r1),
    direction = +1 if r2 & 1 else -1; a re-spawn uses P_HUSKY+1.  The initial draw uses t = 2^63."""

    INIT_STEP = 1 << 63

    def __init__(self, n, seed=0, env_id_base=0, dt=0.01, dist_thresh=0.2, x_offset=0.08, target_z=0.377,
                 respawn_limit=5.0, tables=None):
        f = np.float32
        self.n, self.seed, self.dt, self.thresh = n, seed, f(dt), f(dist_thresh)
        self.x_offset, self.target_z, self.respawn_limit = f(x_offset), f(target_z), f(respawn_limit)
        self.ids = np.arange(n, dtype=np.uint64) + np.uint64(env_id_base)
        self.tables = tables if tables is not None else landing_tables()
        self.traj, self.s = self._draw(self.INIT_STEP)                              # landing.py:210-212
        self.index = np.zeros(n, dtype=np.int32)                                    # :213
        self.pos = np.zeros((n, 2), dtype=f)
        self.heading = np.zeros(n, dtype=f)
        self.step_count = 0

    def _draw(self, t):
        r0, r1, r2, _ = px.draw(self.seed, self.ids, t, px.P_HUSKY)
        traj = px.mulhi(r0, 3).astype(np.int32)
        scale = np.float32(0.8) + np.float32(0.4) * px.u01(r1)                       # landing.py:211 / :241
        direction = np.where((r2 & np.uint32(1)) != 0, np.float32(1), np.float32(-1)).astype(np.float32)
        return traj, (scale * direction).astype(np.float32)

    def _lookup(self):
        out = np.zeros((self.n, 2), dtype=np.float32)
        for k, tab in enumerate(self.tables):
            m = self.traj == k
            out[m] = np.asarray(tab, dtype=np.float32)[np.minimum(self.index[m], len(tab) - 1)]
        return out * self.s[:, None]

    def step(self, reset=None):
        """One set_husky_actions (landing.py:319-364) + unicycle integration.  Returns (wheels [N,4], target3 [N,3])."""
        f = np.float32
        if reset is not None:                                                      # landing.py:263-270
            far = (np.asarray(reset) != 0) & ((np.abs(self.pos[:, 0]) > self.respawn_limit) |
                                              (np.abs(self.pos[:, 1]) > self.respawn_limit))
            r0, r1, _, _ = px.draw(self.seed, self.ids, self.step_count, px.P_HUSKY + 1)
            nx, ny = f(3.0) * px.u01(r0) + f(-1.5), f(3.0) * px.u01(r1) + f(-1.5)
            self.pos[:, 0] = np.where(far, nx, self.pos[:, 0])
            self.pos[:, 1] = np.where(far, ny, self.pos[:, 1])
            self.heading = np.where(far, f(0), self.heading)
        tgt = self._lookup()                                                       # :326-338
        d = tgt - self.pos
        dist = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])                       # :339
        self.index = (self.index + (dist < self.thresh)).astype(np.int32)           # :342-343
        done = (self.index == NUM_WAYPOINTS) | ((self.traj == 2) & (self.index > 3))   # :235-238
        traj, s = self._draw(self.step_count)
        self.traj = np.where(done, traj, self.traj).astype(np.int32)
        self.s = np.where(done, s, self.s).astype(f)
        self.index = np.where(done, 0, self.index).astype(np.int32)
        tgt = self._lookup()                                                       # :349-358
        wheels = differential_drive(self.pos, tgt, self.heading, (3.0, 1000))       # :362
        right, left = wheels[:, 0], wheels[:, 1]
        v = f(WHEEL_RADIUS) * (right + left) * f(0.5)
        wz = f(WHEEL_RADIUS) * (left - right) / f(WHEEL_BASE)
        self.pos = (self.pos + np.stack([np.cos(self.heading), np.sin(self.heading)], -1) * (v * self.dt)[:, None]).astype(f)
        h = self.heading + wz * self.dt
        two_pi = f(2.0) * f(np.pi)
        self.heading = (h - two_pi * np.floor(h / two_pi)).astype(f)
        self.step_count += 1
        target3 = np.stack([self.pos[:, 0] + self.x_offset, self.pos[:, 1], np.full(self.n, self.target_z, f)], -1).astype(f)
        return wheels, target3
