"""`PVFilterBank`: N position / velocity / accel-bias Kalman filters advanced by one kernel launch.

Batched replacement of N `PVFilter` objects (isaacgymenvs/PVFilter.py:6-110) and of the per-env Python loop that
drives them (isaacgymenvs/tasks/ekf_lee_landed.py:417-444).  Method names and argument meaning follow the reference
class; every argument gains a leading env axis.  State lives in SoA planes ([9,N] and [81,N]) for coalescing.
"""
import ctypes as C

import torch

from ._lib import OzlPvArgs, check, lib, ptr


def _s():
    return torch.cuda.current_stream().cuda_stream


class PVFilterBank:
    def __init__(self, num_envs, acc_var, device="cuda:0"):
        if torch.device(device).type != "cuda":
            raise RuntimeError("ouzelum_b200 filters run on CUDA only (no CPU fallback)")
        self.n, self.device = int(num_envs), torch.device(device)
        self.acc_var = [float(v) for v in torch.as_tensor(acc_var).flatten().tolist()]
        self._x = torch.empty(9, self.n, dtype=torch.float32, device=self.device)
        self._P = torch.empty(81, self.n, dtype=torch.float32, device=self.device)
        self.time = 0
        check(lib.ozl_pv_init(self.n, self._x.data_ptr(), self._P.data_ptr(), _s()))

    # ---- reference-style accessors ---------------------------------------------------------------------
    def get_states(self):
        """[N,9] (reference: 9x1 per object, PVFilter.py:16-17)."""
        return self._x.t()

    def get_covariances(self):
        """[N,9,9]."""
        return self._P.t().reshape(self.n, 9, 9)

    def set_states(self, state, env_ids=None):
        state = state.to(self.device, torch.float32)
        if env_ids is None:
            self._x.copy_(state.reshape(self.n, 9).t())
        else:
            self._x[:, env_ids] = state.reshape(-1, 9).t()

    def set_covariances(self, cov):
        self._P.copy_(cov.to(self.device, torch.float32).reshape(self.n, 81).t())

    def reset_states(self, root_states, reset_flags=None):
        """x[flagged] = [pos, vel, 0]  (tasks/ekf_lee_landed.py:353-358)."""
        root = root_states.to(self.device, torch.float32).contiguous()
        check(lib.ozl_pv_reset(self.n, self._x.data_ptr(), ptr(reset_flags), root.data_ptr(), _s()))

    # ---- filter steps ------------------------------------------------------------------------------------
    def step(self, accels=None, orientation=None, dt=0.02, flip_Qw=True, gps_data=None, gps_var=None, gps_mask=None,
             vel_data=None, vel_var=None, vel_mask=None, trigger=None, iter_base=0, vel_var_follows_reference=True):
        """Fused prediction_step + gated position fix + gated velocity fix (one launch).

        trigger = (pos_period, pos_phase, vel_period, vel_phase): when a mask is None, env i takes the fix iff
        (iter_base + i) % period == phase -- the reference's shared counters (ekf_lee_landed.py:425-440) are (7,6,3,0)
        with iter_base = step * N.  `vel_var_follows_reference`: the reference's velocity fix uses R = 0 unless *gps_var*
        is given (PVFilter.py:76-79); set False to use vel_var as written."""
        a = OzlPvArgs()
        a.n, a.x9xN, a.P81xN = self.n, self._x.data_ptr(), self._P.data_ptr()
        keep = []

        def dev(t, shape):
            if t is None:
                return None
            t = torch.as_tensor(t).to(self.device, torch.float32).contiguous()
            if t.dim() == 1 and shape[0] == self.n and t.numel() == shape[1]:
                t = t.expand(self.n, -1).contiguous()
            assert tuple(t.shape) == shape, (tuple(t.shape), shape)
            keep.append(t)
            return t.data_ptr()

        a.do_predict = 1 if accels is not None else 0
        a.accel3, a.quat4 = dev(accels, (self.n, 3)), dev(orientation, (self.n, 4))
        a.pos_meas3, a.vel_meas3 = dev(gps_data, (self.n, 3)), dev(vel_data, (self.n, 3))
        for name, m in (("pos_mask", gps_mask), ("vel_mask", vel_mask)):
            if m is not None:
                m = m.to(self.device).to(torch.uint8).contiguous()
                keep.append(m)
                setattr(a, name, m.data_ptr())
        a.dt = float(dt)
        a.acc_var[:] = self.acc_var
        if gps_var is not None:
            a.pos_var[:] = [float(v) for v in torch.as_tensor(gps_var).flatten().tolist()]
            a.pos_var_given = 1
        if vel_var is not None and (not vel_var_follows_reference or gps_var is not None):
            a.vel_var[:] = [float(v) for v in torch.as_tensor(vel_var).flatten().tolist()]
            a.vel_var_given = 1
        a.flip_qw = 1 if flip_Qw else 0
        if trigger is not None:
            a.pos_period, a.pos_phase, a.vel_period, a.vel_phase = [int(v) for v in trigger]
        else:
            a.pos_period = a.vel_period = 1            # no rule given: every env takes a supplied measurement
        a.iter_base = int(iter_base)
        check(lib.ozl_pv_step(C.byref(a), _s()))
        self._keep = keep                                  # inputs may be temporaries: keep them alive past the launch
        # (same-stream allocator reuse is stream-ordered, so no host synchronisation is needed)

    def prediction_step(self, accels, orientation, dt=0.02, sim_time=0, flip_Qw=True):
        """PVFilter.py:25-64, batched."""
        self.time = sim_time
        self.step(accels=accels, orientation=orientation, dt=dt, flip_Qw=flip_Qw)

    def correction_step(self, gps_data=None, gps_var=None, vel_data=None, vel_var=None, mask=None):
        """PVFilter.py:67-110, batched; velocity fix first, then position fix, as in the reference method."""
        if vel_data is not None:
            self.step(vel_data=vel_data, vel_var=vel_var, gps_var=gps_var, vel_mask=mask)
        if gps_data is not None:
            self.step(gps_data=gps_data, gps_var=gps_var, gps_mask=mask)
