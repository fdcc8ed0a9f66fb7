"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11 "Parallel random numbers: as easy as
1, 2, 3") in numpy -- TEST INFRASTRUCTURE (oracle for ouzelum_b200/csrc/philox.cuh).

The reference draws from torch's global generator in a data-dependent order (SURVEY.md 8a row R:
isaacgymenvs/tasks/ouzelum.py:180-216), which no per-env kernel can reproduce; the framework
instead defines every draw as a pure function of (seed, global env id, global step index,
purpose).  This file is the CPU statement of that function; the CUDA kernels must match it
bit-for-bit (tests/test_philox.py, tests/test_gpu_step.py).
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

# "purpose" words (counter word 3) -- must match ouzelum_b200/csrc/philox.cuh
P_TARGET = 0      # r0,r1 -> target x,y ; r2 -> target z
P_SPAWN = 1       # r0,r1,r2 -> spawn x,y,z offsets
P_FAULT = 2       # r0 -> rotor id ; r1 -> onset step ; r2 -> effectiveness
P_DR0 = 3         # r0..r3 -> mass, Ixx, Iyy, Izz scalings
P_DR1 = 4         # r0,r1,r2 -> arm, thrust-scale, yaw_km
P_DR2 = 24        # second uniforms of gaussian draws (mass, Ixx, Iyy, Izz)
P_DR3 = 25        #                                   (arm, thrust-scale, yaw_km)
P_QDOF0 = 5       # Quadcopter task: initial DOF positions 0..3
P_QDOF1 = 6       #                  initial DOF positions 4..7
P_OBSNOISE = 8    # +0..+3 : 13 per-element sensor-noise uniforms
P_FLICKER = 12    # global (env word = 0xFFFFFFFF) blackout draw
P_ACTION = 16     # synthetic rollout actions (mode B)
P_HUSKY = 20      # waypoint-trajectory re-randomisation (traj id, scale, direction)
GLOBAL_ENV = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All args broadcastable uint32 arrays/ints; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(x, dtype=np.uint64) & MASK32 for x in (c0, c1, c2, c3)])
    c0, c1, c2, c3 = c0.copy(), c1.copy(), c2.copy(), c3.copy()
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def draw(seed, env_ids, step, purpose):
    """The framework's draw function: counter = (env id, step lo, step hi, purpose), key = seed."""
    seed = int(seed)
    step = int(step)
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    return philox4x32_10(env_ids, step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFFFF, purpose,
                         seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def u01(r):
    """uint32 -> float32 uniform in [0,1): top 24 bits * 2^-24 (exact in float32)."""
    return ((np.asarray(r, dtype=np.uint32) >> np.uint32(8)).astype(np.float32)) * np.float32(2.0 ** -24)


def mulhi(r, n):
    """floor(r * n / 2^32): unbiased-enough integer in [0, n) from a uint32 (CUDA __umulhi)."""
    return ((np.asarray(r, dtype=np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)
