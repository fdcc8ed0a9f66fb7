"""Observation / action noise lambdas of the reference's domain randomisation -- TEST INFRASTRUCTURE (oracle for
`noise_lambda_kernel`, ouzelum_b200/csrc/companions.cu, and for `VecTask.apply_randomizations`).

Follows isaacgymenvs/tasks/base/vec_task.py:576-646:
  * schedule (:583-590): linear -> min(last_step, sched_step) / sched_step, constant -> 0 before sched_step and 1 after, none -> 1
  * gaussian (:592-617): additive -> mu, var, mu_corr, var_corr all times the scaling; scaling -> var * s, mu * s + (1 - s) (same for corr);
        lambda(x) = op(x, (corr * var_corr + mu_corr) + randn * var + mu)
  * uniform (:619-644): additive -> lo, hi, lo_corr, hi_corr times the scaling; scaling -> v * s + (1 - s) for all four;
        lambda(x) = op(x, (corr * (hi_corr - lo_corr) + lo_corr) + rand * (hi - lo) + lo)
  * `corr` = randn_like(x), drawn when the lambda first runs after a randomisation event and kept until the next event (:612-616).
The reference draws from torch's global generator; here every sample is a pure function of (seed, env, step | event, element) -- the
counter RNG of oracle/philox.py -- and the kernel must reproduce the same numbers.
"""
import numpy as np

from . import philox as px

P_LAMBDA, P_LAMBDA_CORR = 32, 64
GAUSSIAN, UNIFORM = 3, 1
SCALING, ADDITIVE = 0, 1


def schedule_scaling(sched_type, sched_step, last_step):
    """vec_task.py:583-590."""
    if sched_type == "linear":
        return 1.0 / sched_step * min(last_step, sched_step)
    if sched_type == "constant":
        return 0 if last_step < sched_step else 1
    return 1


def scheduled_params(p, last_step):
    """One `observations` / `actions` block of dr_params -> (distribution, operation, a, b, a_corr, b_corr) after the schedule."""
    dist, op = p["distribution"], p["operation"]
    s = schedule_scaling(p.get("schedule"), p.get("schedule_steps"), last_step) if "schedule" in p else 1
    a, b = p["range"]
    ac, bc = p.get("range_correlated", [0.0, 0.0])
    if dist == "gaussian":
        if op == "additive":
            a, b, ac, bc = a * s, b * s, ac * s, bc * s
        else:
            b, a = b * s, a * s + 1.0 * (1.0 - s)
            bc, ac = bc * s, ac * s + 1.0 * (1.0 - s)
    elif dist == "uniform":
        if op == "additive":
            a, b, ac, bc = a * s, b * s, ac * s, bc * s
        else:
            a, b = a * s + 1.0 * (1.0 - s), b * s + 1.0 * (1.0 - s)
            ac, bc = ac * s + 1.0 * (1.0 - s), bc * s + 1.0 * (1.0 - s)
    else:
        raise ValueError(dist)
    return (GAUSSIAN if dist == "gaussian" else UNIFORM, ADDITIVE if op == "additive" else SCALING,
            np.float32(a), np.float32(b), np.float32(ac), np.float32(bc))


def _box_muller4(r):
    k, two_pi = 5.9604644775390625e-08, 6.283185307179586
    u1 = ((r[0] >> np.uint32(8)).astype(np.float64) + 1.0) * k
    u2 = (r[1] >> np.uint32(8)).astype(np.float64) * k
    u3 = ((r[2] >> np.uint32(8)).astype(np.float64) + 1.0) * k
    u4 = (r[3] >> np.uint32(8)).astype(np.float64) * k
    ra, rb = np.sqrt(-2.0 * np.log(u1)), np.sqrt(-2.0 * np.log(u3))
    f = np.float32
    return [(ra * np.cos(two_pi * u2)).astype(f), (ra * np.sin(two_pi * u2)).astype(f),
            (rb * np.cos(two_pi * u4)).astype(f), (rb * np.sin(two_pi * u4)).astype(f)]


def noise_lambda(x, spec, seed, step, corr_epoch, env_id_base=0, which=0, clip=0.0):
    """x [n,width] float32 -> noised copy; spec = scheduled_params(...)."""
    dist, op, a, b, ac, bc = spec
    f = np.float32
    x = np.asarray(x, dtype=f)
    n, width = x.shape
    ids = np.arange(n, dtype=np.uint64) + np.uint64(env_id_base)
    gauss = dist == GAUSSIAN
    corr_scale, scale = (bc, b) if gauss else (f(bc - ac), f(b - a))
    out = x.copy()
    for j0 in range(0, width, 4):
        g = (j0 >> 2) + 16 * which
        r = px.draw(seed, ids, step, P_LAMBDA + g)
        rc = px.draw(seed, ids, corr_epoch, P_LAMBDA_CORR + g)
        zc = _box_muller4(rc)
        zf = _box_muller4(r) if gauss else [px.u01(v) for v in r]
        for j in range(min(4, width - j0)):
            corr = (zc[j] * corr_scale + ac).astype(f)
            noise = ((corr + (zf[j] * scale).astype(f)).astype(f) + a).astype(f)
            v = out[:, j0 + j]
            v = (v + noise).astype(f) if op == ADDITIVE else (v * noise).astype(f)
            if clip > 0:
                v = np.minimum(np.maximum(v, f(-clip)), f(clip))
            out[:, j0 + j] = v
    return out
