"""CPU oracle of the Lee geometric controllers (kernel K4) -- TEST INFRASTRUCTURE.

Component-wise numpy restatement (float32 or float64) of
  isaacgymenvs/controllers/controller.py:45-48            Controller.__call__ (input scaling + dispatch)
  isaacgymenvs/controllers/position_control.py:19-109     LeePositionController.__call__
  isaacgymenvs/controllers/velocity_control.py:17-112     LeeVelocityController.__call__
  isaacgymenvs/controllers/attitude_control.py:17-78      LeeAttitudeContoller.__call__
  isaacgymenvs/controllers/rotation_conversions.py:36-64  quaternion_to_matrix (wxyz, two_s = 2/|q|^2)
  isaacgymenvs/controllers/rotation_conversions.py:149-171,216-255  euler_angles_to_matrix / matrix_to_euler_angles("ZYX")
  isaacgymenvs/controllers/math_control.py:10-16          compute_vee_map
  isaacgymenvs/controllers/control_config.py:13-18        gains
Pinned against the reference's own classes executed on CPU (tests/golden/lee_*.npz, made by
tests/golden/make_golden.py).  The reference writes these as batched torch ops (bmm of 3x3); here every
matrix product is spelled out per component, which is also the form the CUDA kernel uses.

Known reference quirk kept out of scope: `torch.cross(b2_c, b3_c)` without `dim` (position_control.py:60)
picks dim 0 when the batch size is exactly 3.
"""
import numpy as np

KP = (0.8, 0.8, 1.0)      # control_config.py:14
KV = (0.5, 0.5, 0.4)      # control_config.py:15
KR = (3.0, 3.0, 1.0)      # control_config.py:16
KOMEGA = (0.5, 0.5, 1.20)  # control_config.py:17
SCALE_INPUT = (1.0, 1.0, 1.0, 1.0)   # control_config.py:18
PI = 3.14159265358979323846          # literal used at position_control.py:90-92

POSITION, VELOCITY, ATTITUDE = 0, 1, 2


def _quat_wxyz_to_R(r, i, j, k):
    """rotation_conversions.py:36-64."""
    two_s = 2.0 / (((r * r + i * i) + j * j) + k * k)
    return [[1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r)],
            [two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r)],
            [two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)]]


def _euler_zyx(R):
    """matrix_to_euler_angles(R, "ZYX")[:, [2, 1, 0]] -> (roll, pitch, yaw)   rotation_conversions.py:216-255."""
    yaw = np.arctan2(R[1][0], R[0][0])
    pitch = np.arcsin(-R[2][0])
    roll = np.arctan2(R[2][1], R[2][2])
    return roll, pitch, yaw


def _R_from_zyx(yaw, pitch, roll):
    """euler_angles_to_matrix((yaw, pitch, roll), "ZYX") = Rz(yaw) Ry(pitch) Rx(roll)   rotation_conversions.py:149-171."""
    cz, sz, cy, sy, cx, sx = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    zero, one = np.zeros_like(cz), np.ones_like(cz)
    Rz = [[cz, -sz, zero], [sz, cz, zero], [zero, zero, one]]
    Ry = [[cy, zero, sy], [zero, one, zero], [-sy, zero, cy]]
    Rx = [[one, zero, zero], [zero, cx, -sx], [zero, sx, cx]]
    return _mm(_mm(Rz, Ry), Rx)


def _mm(A, B):
    return [[(A[i][0] * B[0][j] + A[i][1] * B[1][j]) + A[i][2] * B[2][j] for j in range(3)] for i in range(3)]


def _T(A):
    return [[A[j][i] for j in range(3)] for i in range(3)]


def _mv(A, v):
    return [(A[i][0] * v[0] + A[i][1] * v[1]) + A[i][2] * v[2] for i in range(3)]


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _norm(a):
    return np.sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2])


def _attitude_tail(R, Rd, roll, pitch, yaw_rate, omega_w, kR, kO):
    """Shared tail: rotation error, body-rate setpoint, torque (position_control.py:66-109)."""
    Rt, Rdt = _T(R), _T(Rd)
    A, B = _mm(Rdt, R), _mm(Rt, Rd)
    M = [[A[i][j] - B[i][j] for j in range(3)] for i in range(3)]
    e_R = [0.5 * (-M[1][2]), 0.5 * M[0][2], 0.5 * (-M[0][1])]          # math_control.py:10-16
    s_p, c_p, s_r, c_r = np.sin(pitch), np.cos(pitch), np.sin(roll), np.cos(roll)
    w_des_b = [-s_p * yaw_rate, (s_r * c_p) * yaw_rate, (c_r * c_p) * yaw_rate]
    desired = _mv(Rt, _mv(Rd, w_des_b))
    actual = _mv(Rt, omega_w)
    return [-kR[j] * e_R[j] - kO[j] * (actual[j] - desired[j]) for j in range(3)]


def lee_control(state, command, mode=POSITION, kP=KP, kV=KV, kR=KR, kO=KOMEGA, scale=SCALE_INPUT, dtype=np.float32):
    """state [N,13] (pos, quat xyzw, linvel, angvel world), command [N,4] -> (thrust [N], torque [N,3])."""
    f = dtype
    state = np.asarray(state, dtype=f)
    cmd = np.asarray(command, dtype=f) * np.asarray(scale, dtype=f)           # controller.py:47
    kP, kV, kR, kO = (np.asarray(g, dtype=f) for g in (kP, kV, kR, kO))
    p = [state[:, j] for j in range(3)]
    v = [state[:, 7 + j] for j in range(3)]
    w = [state[:, 10 + j] for j in range(3)]
    R = _quat_wxyz_to_R(state[:, 6], state[:, 3], state[:, 4], state[:, 5])
    roll, pitch, yaw = _euler_zyx(R)
    b3col = [R[0][2], R[1][2], R[2][2]]

    if mode == POSITION:
        acc = [kP[j] * (cmd[:, j] - p[j]) - kV[j] * v[j] for j in range(3)]      # position_control.py:39-41
        acc[2] = acc[2] + f(1)
        thrust = (acc[0] * b3col[0] + acc[1] * b3col[1]) + acc[2] * b3col[2]       # :44
        n = _norm(acc)
        b3 = [a / n for a in acc]                                                  # :47
        tmp = [np.cos(yaw), np.sin(yaw), np.zeros_like(yaw)]                       # :50-52
        b2 = _cross(b3, tmp)
        n2 = _norm(b2)
        b2 = [a / n2 for a in b2]
        b1 = _cross(b2, b3)                                                        # :60
        Rd = [[b1[i], b2[i], b3[i]] for i in range(3)]
        yr = np.remainder(cmd[:, 3] - yaw, f(PI * 2.0))                            # :90
        yr = np.where(yr > f(PI), yr - f(PI * 2.0), yr)                            # :92
    elif mode == VELOCITY:
        zero = np.zeros_like(yaw)
        Rv = _R_from_zyx(yaw, zero, zero)                                          # velocity_control.py:32-38
        vv = _mv(_T(Rv), v)
        acc = [kV[j] * (cmd[:, j] - vv[j]) for j in range(3)]                      # :46-48
        acc[2] = acc[2] + f(1)
        thrust = (acc[0] * b3col[0] + acc[1] * b3col[1]) + acc[2] * b3col[2]
        pitch_sp = np.arctan2(acc[0], acc[2])                                      # :58
        roll_sp = np.arctan2(-acc[1], np.sqrt(acc[2] ** 2 + acc[0] ** 2))          # :59-60
        Rd = _R_from_zyx(yaw, pitch_sp, roll_sp)                                   # :70-71
        yr = cmd[:, 3]                                                             # :93
    elif mode == ATTITUDE:
        Rd = _R_from_zyx(yaw, cmd[:, 2], cmd[:, 1])                                # attitude_control.py:33-35,58-59
        thrust = cmd[:, 0] + f(1)                                                  # :78
        yr = cmd[:, 3]                                                             # :53
    else:
        raise ValueError("Invalid controller name: {}".format(mode))              # controller.py:34
    torque = _attitude_tail(R, Rd, roll, pitch, yr, w, kR, kO)
    return np.asarray(thrust, dtype=f), np.stack(torque, -1).astype(f)
