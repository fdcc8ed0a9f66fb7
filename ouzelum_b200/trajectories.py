"""Waypoint tables and the batched ground-vehicle follower (kernel K5).

`lemniscate / circle / square` produce the reference's tables (isaacgymenvs/utils/trajectories.py:5-60) -- they are
init-only, so they are evaluated once on the host with the same float32 / float64 arithmetic the reference uses and
uploaded.  `HuskyFollower` replaces the two O(N) Python loops of `Landing.set_husky_actions`
(isaacgymenvs/tasks/landing.py:319-364) with one kernel launch per step.
"""
import ctypes as C
import math

import torch

from ._lib import OzlHuskyArgs, check, lib, ptr


def lemniscate(a: float = math.sqrt(2), num_points: int = 200):
    th = torch.linspace(-math.pi / 2, 3 * math.pi / 2, num_points)
    s, c = torch.sin(th), torch.cos(th)
    den = s ** 2 + 1
    return torch.stack((a * c / den, a * c * s / den), dim=1)


def circle(r: float = math.sqrt(2), num_points: int = 200):
    step = 360 / num_points
    ang = [math.radians(i * step) for i in range(num_points)]
    return torch.tensor([(r * math.cos(t), r * math.sin(t)) for t in ang])


def square(side_length: float = 5, num_points: int = 8):
    if num_points < 4:
        raise ValueError("A square needs at least 4 waypoints.")
    k = num_points // 4
    d = side_length / (k - 1)
    pts = [(i * d, 0) for i in range(k)] + [(side_length, i * d) for i in range(1, k)] + \
          [(side_length - i * d, side_length) for i in range(1, k)] + [(0, side_length - i * d) for i in range(1, k - 1)]
    return -(torch.tensor(pts) - (side_length / 2))


def landing_tables(device):
    """[204,2] float32: lemniscate(a=4,100) | circle(r=2,100) | square(4,8)   (landing.py:108-112)."""
    return torch.cat([lemniscate(4, 100), circle(2, 100), square(4, 8)], 0).to(device=device, dtype=torch.float32).contiguous()


class HuskyFollower:
    """N ground vehicles, each following a randomly chosen / scaled / mirrored waypoint trajectory."""

    def __init__(self, num_envs, device="cuda:0", seed=0, env_id_base=0, dt=0.01, dist_thresh=0.2, x_offset=0.08,
                 target_z=0.377, env_spacing=2.5):
        if torch.device(device).type != "cuda":
            raise RuntimeError("ouzelum_b200.HuskyFollower runs on CUDA only (no CPU fallback)")
        self.n, self.device = int(num_envs), torch.device(device)
        self.tables = landing_tables(self.device)
        self.pose = torch.zeros(self.n, 4, dtype=torch.float32, device=self.device)      # x, y, heading, scale*direction
        self.idx = torch.zeros(self.n, 2, dtype=torch.int32, device=self.device)         # trajectory id, waypoint index
        self.wheels = torch.zeros(self.n, 4, dtype=torch.float32, device=self.device)
        self.target = torch.zeros(self.n, 3, dtype=torch.float32, device=self.device)
        self.step_count = 0
        a = self._a = OzlHuskyArgs()
        a.n, a.pose4, a.idx2, a.tables204x2 = self.n, self.pose.data_ptr(), self.idx.data_ptr(), self.tables.data_ptr()
        a.wheels4, a.target3 = self.wheels.data_ptr(), self.target.data_ptr()
        a.seed, a.env_id_base, a.dt, a.dist_thresh = int(seed), int(env_id_base), float(dt), float(dist_thresh)
        a.kp_lin, a.kp_ang, a.ang_thresh = 3.0, 1000.0, 0.005                            # landing.py:362, controllers.py:16
        a.x_offset, a.target_z, a.respawn_limit = float(x_offset), float(target_z), 2.0 * float(env_spacing)
        check(lib.ozl_husky_init(C.byref(a), torch.cuda.current_stream().cuda_stream))

    # reference-style views (landing.py:84-88,209-213)
    @property
    def husky_positions(self):
        return self.pose[:, 0:2]

    @property
    def husky_trajectories(self):
        return self.idx[:, 0]

    @property
    def target_indices(self):
        return self.idx[:, 1]

    def follow_step_counter(self, device_ptr):
        """Read the step index from a device counter (ozl_step_counter_ptr) instead of the host-side count: makes the
        launch free of host-changing arguments (CUDA-graph capturable)."""
        self._a.step_ptr = device_ptr

    def step(self, reset_buf=None):
        """Advance every vehicle one control step; returns the landing target [N,3] riding on it."""
        a = self._a
        a.reset = ptr(reset_buf)
        a.step = self.step_count
        check(lib.ozl_husky_step(C.byref(a), torch.cuda.current_stream().cuda_stream))
        self.step_count += 1
        return self.target
