"""CPU oracle of the batched position / velocity / accel-bias Kalman filter (kernel K3) -- TEST INFRASTRUCTURE.

numpy restatement, batched over envs (leading axis), of
  isaacgymenvs/PVFilter.py:7-14     state x = [p(3), v(3), b_a(3)], P = 1000 I9
  isaacgymenvs/PVFilter.py:25-64    prediction_step
  isaacgymenvs/PVFilter.py:67-110   correction_step
  isaacgymenvs/PVFilter.py:113-142  quaternion_to_matrix (normalises first)
The reference keeps one Python object per env and loops (tasks/ekf_lee_landed.py:417-444); its quirks are
reproduced on purpose:
  * R_body_to_nav is the TRANSPOSE of quaternion_to_matrix (PVFilter.py:33-35)
  * F[3:6,3:6] = R (replaces the identity: the velocity block is rotated every step)   (:51)
  * G = F[0:6,6:9] (the bias columns reused as the input matrix)                         (:54-55)
  * in the velocity fix the measurement noise is zero unless *gps_var* is given          (:76-79)
Pinned against the reference class executed on CPU (tests/golden/pvfilter.npz).
"""
import numpy as np


def quat_to_matrix_wxyz(q):
    """PVFilter.py:113-142: q/|q| then the PyTorch3D formula."""
    q = q / np.sqrt(np.sum(q * q, axis=-1, keepdims=True))
    r, i, j, k = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    two_s = 2.0 / np.sum(q * q, axis=-1)
    o = np.stack([1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                  two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                  two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)], -1)
    return o.reshape(q.shape[:-1] + (3, 3))


class PVFilterBank:
    """N independent filters, state [N,9], cov [N,9,9]."""

    def __init__(self, n, acc_var, dtype=np.float32):
        self.f = dtype
        self.n = n
        self.state = np.zeros((n, 9), dtype=dtype)                                   # PVFilter.py:11
        self.cov = np.broadcast_to(np.eye(9, dtype=dtype) * dtype(1000), (n, 9, 9)).copy()   # :12
        self.acc_var = np.diag(np.asarray(acc_var, dtype=dtype))                     # :13

    def set_states(self, idx, state):
        self.state[idx] = state

    def prediction_step(self, accels, orientation, dt, flip_Qw=True, mask=None):
        """accels [N,3], orientation [N,4] (xyzw if flip_Qw else wxyz).  PVFilter.py:25-64."""
        f = self.f
        accels = np.asarray(accels, dtype=f)
        q = np.asarray(orientation, dtype=f)
        q = q[:, [3, 0, 1, 2]] if flip_Qw else q
        Rbn = np.swapaxes(quat_to_matrix_wxyz(q), -1, -2)                            # .T  (:33-35)
        dt = f(dt)
        F = np.broadcast_to(np.eye(9, dtype=f), (self.n, 9, 9)).copy()
        F[:, 0:3, 3:6] = Rbn * dt
        F[:, 0:3, 6:9] = Rbn * (dt ** 2) * f(0.5)
        F[:, 3:6, 3:6] = Rbn
        F[:, 3:6, 6:9] = Rbn * dt
        G = np.zeros((self.n, 9, 3), dtype=f)
        G[:, 0:6, :] = F[:, 0:6, 6:9]
        u = accels - self.state[:, 6:9]
        new_state = np.einsum("nij,nj->ni", F, self.state) + np.einsum("nij,nj->ni", G, u)
        new_cov = F @ self.cov @ np.swapaxes(F, -1, -2) + G @ self.acc_var @ np.swapaxes(G, -1, -2)
        if mask is None:
            self.state, self.cov = new_state.astype(f), new_cov.astype(f)
        else:
            self.state[mask], self.cov[mask] = new_state[mask], new_cov[mask]

    def _correct(self, z, lo, R, mask):
        f = self.f
        H = slice(lo, lo + 3)
        S = self.cov[:, H, H] + R
        K = self.cov[:, :, H] @ np.linalg.inv(S.astype(np.float64)).astype(f)       # :82 / :102
        innov = np.asarray(z, dtype=f) - self.state[:, H]
        new_state = self.state + np.einsum("nij,nj->ni", K, innov)
        IKH = np.broadcast_to(np.eye(9, dtype=f), (self.n, 9, 9)).copy()
        IKH[:, :, H] = IKH[:, :, H] - K
        new_cov = IKH @ self.cov
        if mask is None:
            mask = np.ones(self.n, dtype=bool)
        self.state[mask], self.cov[mask] = new_state[mask].astype(f), new_cov[mask].astype(f)

    def correction_step(self, gps_data=None, gps_var=None, vel_data=None, vel_var=None, mask=None):
        """PVFilter.py:67-110.  `mask` selects which envs are corrected (the reference decides per env)."""
        f = self.f
        if vel_data is not None:
            R = np.zeros((3, 3), dtype=f) if gps_var is None else np.diag(np.asarray(vel_var, dtype=f))   # :76-79 (sic)
            self._correct(vel_data, 3, R, mask)
        if gps_data is not None:
            R = np.zeros((3, 3), dtype=f) if gps_var is None else np.diag(np.asarray(gps_var, dtype=f))   # :96-99
            self._correct(gps_data, 0, R, mask)
