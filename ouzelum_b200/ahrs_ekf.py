"""`EKFBank`: N attitude EKFs (float64) advanced by one kernel launch.

Batched replacement of N `EKF` objects (isaacgymenvs/ahrs_ekf.py:879-1337, a modified copy of ahrs.filters.EKF) on the
only branch the quadcopter tasks execute: `update(q, gyr, ang, acc)` with a direct quaternion measurement
(isaacgymenvs/tasks/ekf_lee_landed.py:378-391).  No D2H copy, no Python loop, no NumPy.
"""
import torch

from ._lib import check, lib, ptr


def _s():
    return torch.cuda.current_stream().cuda_stream


class EKFBank:
    def __init__(self, num_envs, frequency=100.0, device="cuda:0", **kwargs):
        if torch.device(device).type != "cuda":
            raise RuntimeError("ouzelum_b200 filters run on CUDA only (no CPU fallback)")
        self.n, self.device = int(num_envs), torch.device(device)
        self.frequency = frequency
        self.Dt = kwargs.get("Dt", 1.0 / self.frequency)                      # ahrs_ekf.py:993
        noises = list(kwargs.get("noises", [0.3 ** 2, 0.5 ** 2, 0.8 ** 2]))   # ahrs_ekf.py:1004
        if "var_gyr" in kwargs:
            noises[0] = kwargs["var_gyr"]
        self.g_noise = float(noises[0])
        self._q = torch.empty(4, self.n, dtype=torch.float64, device=self.device)
        self._P = torch.empty(16, self.n, dtype=torch.float64, device=self.device)
        check(lib.ozl_ekf_init(self.n, self._q.data_ptr(), self._P.data_ptr(), _s()))

    @property
    def Q_state(self):
        """[N,4] wxyz (reference: self.Q_state, ekf_lee_landed.py:143)."""
        return self._q.t()

    @property
    def P(self):
        """[N,4,4]."""
        return self._P.t().reshape(self.n, 4, 4)

    def set_q_from_root_quats(self, root_quats_xyzw, flags=None):
        """Q_state[flagged] = root_quats[flagged][:, [3,0,1,2]]  (ekf_lee_landed.py:349-352); flags None = all."""
        q = root_quats_xyzw.to(self.device, torch.float32).contiguous()
        check(lib.ozl_ekf_set_q(self.n, self._q.data_ptr(), q.data_ptr(), ptr(flags), _s()))

    def set_state(self, q_wxyz, P=None):
        self._q.copy_(q_wxyz.to(self.device, torch.float64).reshape(self.n, 4).t())
        if P is not None:
            self._P.copy_(P.to(self.device, torch.float64).reshape(self.n, 16).t())

    def update(self, gyr, ang, acc=None, ang_xyzw=False, check_norm=False, q_f32_out=None):
        """EKF.update for every env: gyr [N,3] f32, ang [N,4] f32 (wxyz, or xyzw with ang_xyzw=True).  `acc` is accepted
        for signature parity and unused (it only feeds dead code on this branch).  Returns Q_state [N,4]."""
        if check_norm:                                                       # ahrs_ekf.py:1301-1302 (host sync: debug only)
            nrm = self._q.norm(dim=0)
            if not torch.allclose(nrm, torch.ones_like(nrm)):
                raise ValueError("A-priori quaternion must have a norm equal to 1.")
        g = gyr.to(self.device, torch.float32).contiguous()
        a = ang.to(self.device, torch.float32).contiguous()
        check(lib.ozl_ekf_update(self.n, self._q.data_ptr(), self._P.data_ptr(), g.data_ptr(), a.data_ptr(),
                                 1 if ang_xyzw else 0, float(self.Dt), self.g_noise, ptr(q_f32_out), _s()))
        return self.Q_state
