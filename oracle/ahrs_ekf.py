"""CPU oracle of the batched attitude EKF (kernel K2) -- TEST INFRASTRUCTURE.

numpy float64 restatement, batched over envs, of the branch of the reference's modified `ahrs` EKF that the
quadcopter tasks execute (a direct quaternion measurement `ang` is always supplied, tasks/ekf_lee_landed.py:378-391):
  isaacgymenvs/ahrs_ekf.py:982-1012    __init__ / noises: P = I4, g_noise = 0.3^2, Dt = 1/frequency
  isaacgymenvs/ahrs_ekf.py:1072-1106   Omega
  isaacgymenvs/ahrs_ekf.py:1108-1133   f      q_t = (I + Dt/2 Omega(g)) q
  isaacgymenvs/ahrs_ekf.py:1135-1158   dfdq   F = I + Omega(Dt/2 g)
  isaacgymenvs/ahrs_ekf.py:1280-1337   update (`ang` branch :1329-1332: v = ang - q_t, H = I, S = P_t + 1e-7 I)
Third-party arithmetic: `skew` comes from the un-pinned PyPI package `ahrs` (setup.py:19, not in the reference
tree); its published definition skew(x) = [[0,-x2,x1],[x2,0,-x0],[-x1,x0,0]] is restated here.  Pinned against the
reference's own `EKF.update` source executed with that one helper stubbed (tests/golden/ekf.npz).
The accelerometer enters only through `a /= |a|` (:1309), whose result is unused on this branch; a zero
accelerometer vector therefore yields NaN in the reference's `self.z` only, never in q -- nothing to reproduce.
"""
import numpy as np

G_NOISE = 0.3 ** 2        # ahrs_ekf.py:1004
S_EPS = 0.0000001         # ahrs_ekf.py:1332


def omega(x):
    """ahrs_ekf.py:1100-1106, batched: x [N,3] -> [N,4,4]."""
    z = np.zeros_like(x[:, 0])
    return np.stack([np.stack([z, -x[:, 0], -x[:, 1], -x[:, 2]], -1),
                     np.stack([x[:, 0], z, x[:, 2], -x[:, 1]], -1),
                     np.stack([x[:, 1], -x[:, 2], z, x[:, 0]], -1),
                     np.stack([x[:, 2], x[:, 1], -x[:, 0], z], -1)], -2)


def skew(x):
    z = np.zeros_like(x[:, 0])
    return np.stack([np.stack([z, -x[:, 2], x[:, 1]], -1),
                     np.stack([x[:, 2], z, -x[:, 0]], -1),
                     np.stack([-x[:, 1], x[:, 0], z], -1)], -2)


class EKFBank:
    """N independent filters: P [N,4,4] float64, Dt = 1/frequency."""

    def __init__(self, n, frequency=100.0, dtype=np.float64):
        self.n, self.f = n, dtype
        self.Dt = dtype(1.0 / frequency)                                    # :993
        self.P = np.broadcast_to(np.identity(4, dtype=dtype), (n, 4, 4)).copy()   # :995
        self.g_noise = dtype(G_NOISE)

    def update(self, q, gyr, ang, acc=None):
        """q [N,4] wxyz (|q| ~ 1 else ValueError, :1301-1302), gyr [N,3], ang [N,4] wxyz -> q [N,4]."""
        f = self.f
        q, g, ang = np.asarray(q, dtype=f), np.asarray(gyr, dtype=f), np.asarray(ang, dtype=f)
        if not np.all(np.isclose(np.linalg.norm(q, axis=-1), 1.0)):
            raise ValueError("A-priori quaternion must have a norm equal to 1.")
        I4 = np.identity(4, dtype=f)
        q_t = np.einsum("nij,nj->ni", I4 + f(0.5) * self.Dt * omega(g), q)           # :1132-1133
        F = I4 + omega(f(0.5) * self.Dt * g)                                          # :1157-1158
        top = -q[:, None, 1:]                                                         # [-q[1:]]
        W = f(0.5) * self.Dt * np.concatenate([top, q[:, 0, None, None] * np.identity(3, dtype=f) + skew(q[:, 1:])], 1)  # :1320
        Q_t = f(0.5) * self.Dt * self.g_noise * (W @ np.swapaxes(W, -1, -2))          # :1321
        P_t = F @ self.P @ np.swapaxes(F, -1, -2) + Q_t                               # :1322
        v = ang - q_t                                                                 # :1330
        S = P_t + I4 * f(S_EPS)                                                       # :1332
        K = P_t @ np.linalg.inv(S)                                                    # :1333
        self.P = (I4 - K) @ P_t                                                       # :1334
        qn = q_t + np.einsum("nij,nj->ni", K, v)                                      # :1335
        return qn / np.linalg.norm(qn, axis=-1, keepdims=True)                        # :1336
