// Small device helpers shared by the stand-alone E3 glue kernels (companions.cu) and the fused estimator + controller
// kernel (ekf_lee_fused.cu).  Reference: isaacgymenvs/tasks/ekf_lee_landed.py:339-503, isaacgymenvs/utils/POMDP.py:23-42.
#pragma once
#include "philox.cuh"

namespace ozl {

struct FaultCfg {
    int mode;                // 0: truth ; 1..3: OZL_POMDP_*
    float flicker_p, noise_lo, noise_range;
    uint64_t seed, step;
};

// POMDPWrapper.observation on a d <= 4 vector of one env.  Batched calls of the reference draw ONE flicker value for all
// envs (GLOBAL_ENV); the per-env quaternion call draws one per env (ekf_lee_landed.py:383).  `stream` separates the uses.
__device__ __forceinline__ void sensor_fault(const FaultCfg& a, uint32_t genv, uint32_t stream, bool per_env_flicker, float* v,
                                             int d) {
    if (a.mode == 0) return;
    bool black = false;
    if (a.mode == 1 || a.mode == 3) {
        const uint4 r = draw(a.seed, per_env_flicker ? genv : GLOBAL_ENV, a.step, P_FLICKER + (stream << 8));
        black = u01(r.x) <= a.flicker_p;
    }
    uint4 r = make_uint4(0, 0, 0, 0);
    if (a.mode >= 2) r = draw(a.seed, genv, a.step, P_OBSNOISE + (stream << 8));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    for (int j = 0; j < d; ++j) {
        float x = black ? 0.0f : v[j];
        // explicit _rn intrinsics: never contracted to FMA, so the noise factor is bit-identical to the oracle's and to the
        // in-kernel observation noise of quad_step (this TU is compiled with FMA contraction on)
        if (a.mode >= 2) x = __fmul_rn(x, __fadd_rn(__fmul_rn(u01(rr[j]), a.noise_range), a.noise_lo));
        v[j] = x;
    }
}

// Carrot-waypoint logic (ekf_lee_landed.py:458-486).  p = true position, t = target, w = waypoint (in/out).
//   warm-up: waypoint = target
//   after:   if waypoint_dist < 0.5 or > 1.0: waypoint = pos + 0.75 * unit(target + (0,0,0.7) - pos)
//            if target_dist < 0.75:           waypoint = target + (0,0,0.09)
// The reference guards the carrot update with a GLOBAL `(waypoint_dist == 0).any()` (:474); here the guard is per env.
__device__ __forceinline__ void waypoint_update(const float p[3], const float t[3], float w[3], bool warm) {
    if (warm) { w[0] = t[0]; w[1] = t[1]; w[2] = t[2]; }                                    // :461-463
    const float tv[3] = {t[0] - p[0], t[1] - p[1], t[2] - p[2]};
    const float td = sqrtf((tv[0] * tv[0] + tv[1] * tv[1]) + tv[2] * tv[2]);
    const float wv[3] = {w[0] - p[0], w[1] - p[1], w[2] - p[2]};
    const float wd = sqrtf((wv[0] * wv[0] + wv[1] * wv[1]) + wv[2] * wv[2]);
    if (!warm) {
        if ((wd < 0.5f || wd > 1.0f) && wd != 0.0f) {                                       // :473-482
            const float rv[3] = {t[0] - p[0], t[1] - p[1], (t[2] + 0.7f) - p[2]};
            const float rd = sqrtf((rv[0] * rv[0] + rv[1] * rv[1]) + rv[2] * rv[2]);
            for (int j = 0; j < 3; ++j) w[j] = (rv[j] / rd) * 0.75f + p[j];
        }
        if (td < 0.75f) { w[0] = t[0]; w[1] = t[1]; w[2] = t[2] + 0.09f; }                  // :483-486
    }
}

}  // namespace ozl
