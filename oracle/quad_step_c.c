/* C restatement of the fused x500 env step -- TEST INFRASTRUCTURE (second, independent CPU oracle + the multi-threaded
 * CPU baseline of bench.py).  Same float32 operation order as oracle/quad_step.py (QuadStepOracle, rotor-action mode) and
 * as the CUDA kernel; compile with -ffp-contract=off so that nothing is fused.  Reference lines restated:
 *   isaacgymenvs/tasks/base/vec_task.py:313-359, isaacgymenvs/tasks/ouzelum.py:180-332,
 *   isaacgymenvs/utils/torch_jit_utils.py:66-71,198-208, isaacgymenvs/utils/POMDP.py:23-42;
 *   gym.simulate -> the integrator of SURVEY.md 8a row P.
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp -shared -fPIC -Iinclude oracle/quad_step_c.c -lm   (oracle/c_oracle.py;
 * -mfma only lets the EXPLICIT fmaf calls compile to one instruction -- without it glibc's exact software fmaf is used)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "ouzelum_b200.h"

enum { P_TARGET = 0, P_SPAWN = 1, P_FAULT = 2, P_DR0 = 3, P_DR1 = 4, P_OBSNOISE = 8, P_FLICKER = 12, P_DR2 = 24, P_DR3 = 25 };
#define GLOBAL_ENV 0xFFFFFFFFu
#define FAULT_NEVER 0x1FFFFFFF

static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static void draw(uint64_t seed, uint32_t env, uint64_t step, uint32_t purpose, uint32_t out[4]) {
    philox(env, (uint32_t)step, (uint32_t)(step >> 32), purpose, (uint32_t)seed, (uint32_t)(seed >> 32), out);
}
static float u01(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }

/* one domain-randomised parameter (isaacgymenvs/utils/dr_utils.py:71-132); same operations as dr_apply() in quad_env.cuh */
static float dr_apply(const ozl_dr_param* d, float nominal, uint32_t r0, uint32_t r1, uint64_t step) {
    if (d->distribution == OZL_DR_NONE) return nominal;
    float a = d->range[0], b = d->range[1];
    if (d->schedule != OZL_DR_SCHED_NONE) {
        const uint64_t lim = (uint64_t)d->schedule_steps;
        const float inv_steps = 1.0f / (float)d->schedule_steps;
        const float ss = d->schedule == OZL_DR_SCHED_LINEAR ? inv_steps * (float)(step < lim ? step : lim) : (step < lim ? 0.0f : 1.0f);
        const float one_m = 1.0f - ss;
        if (d->operation == OZL_DR_ADDITIVE) { a = a * ss; b = b * ss; }
        else if (d->distribution == OZL_DR_GAUSSIAN) { a = a * ss + one_m; b = b * ss; }
        else { a = a * ss + one_m; b = b * ss + one_m; }
    }
    float smp;
    if (d->distribution == OZL_DR_UNIFORM) smp = a + (b - a) * u01(r0);
    else if (d->distribution == OZL_DR_LOGUNIFORM) { const double la = log((double)a), lb = log((double)b); smp = (float)exp(la + (lb - la) * (double)u01(r0)); }
    else {
        const double u1 = ((double)(r0 >> 8) + 1.0) * 5.9604644775390625e-08, u2 = (double)u01(r1);
        smp = (float)((double)a + (double)b * (sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2)));
    }
    return d->operation == OZL_DR_ADDITIVE ? nominal + smp : nominal * smp;
}

typedef struct {
    float h, hh, hh2, max_angvel2, inv3, half, inv_pi, flicker_p, noise_lo, noise_range;
    float sinc_c1, sinc_c2, cos_c1, cos_c2, cos_c3;
    int nsub;
} derived_t;

/* integrator arithmetic "row P" v2: single IEEE operations and explicit fmaf, in the order of quad_env.cuh */
static void quat_to_R(const float q[4], float R[3][3]) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float wx = x2 * w, wy = y2 * w, wz = z2 * w;
    const float a = fmaf(-y2, y, 1.0f), b = fmaf(-x2, x, 1.0f);
    R[0][0] = fmaf(-z2, z, a);  R[0][1] = fmaf(x2, y, -wz); R[0][2] = fmaf(x2, z, wy);
    R[1][0] = fmaf(x2, y, wz);  R[1][1] = fmaf(-z2, z, b);  R[1][2] = fmaf(y2, z, -wx);
    R[2][0] = fmaf(x2, z, -wy); R[2][1] = fmaf(y2, z, wx);  R[2][2] = fmaf(-y2, y, b);
}
static void mv(float R[3][3], const float v[3], float o[3]) {
    for (int i = 0; i < 3; ++i) o[i] = fmaf(R[i][2], v[2], fmaf(R[i][1], v[1], R[i][0] * v[0]));
}
static void mtv(float R[3][3], const float v[3], float o[3]) {
    for (int i = 0; i < 3; ++i) o[i] = fmaf(R[2][i], v[2], fmaf(R[1][i], v[1], R[0][i] * v[0]));
}
static void cross3(const float a[3], const float b[3], float o[3]) {
    o[0] = fmaf(a[1], b[2], -(a[2] * b[1])); o[1] = fmaf(a[2], b[0], -(a[0] * b[2])); o[2] = fmaf(a[0], b[1], -(a[1] * b[0]));
}

/* state arrays (AoS, caller-owned): root [n,13], thrust [n,4], target [n,3], ep_ret [n], params [n,8] =
 * mass,ixx,iyy,izz,arm,thrust_scale,fault_eff,yaw_km, fault [n,2] = rotor, onset */
void ozl_oracle_step(const ozl_cfg* c, uint64_t step, int64_t n, float* root, float* thrust, float* target, float* ep_ret,
                     float* params, int32_t* fault, const float* actions, float* obs, float* rew, int64_t* reset,
                     int64_t* progress, uint8_t* timeout) {
    derived_t d;
    const double h = (double)c->dt / (double)c->substeps;
    d.h = (float)h; d.hh = (float)(0.5 * h); d.hh2 = d.hh * d.hh;
    d.max_angvel2 = (float)((double)c->max_angvel * (double)c->max_angvel);
    d.inv3 = 1.0f / 3.0f; d.half = 0.5f; d.inv_pi = 1.0f / (float)M_PI;
    d.flicker_p = (c->pomdp_mode == 3) ? 0.1f : c->pomdp_prob;
    { const float lo = (float)(1.0 - (double)c->noise_sigma), hi = (float)(1.0 + (double)c->noise_sigma); d.noise_lo = lo; d.noise_range = hi - lo; }
    d.sinc_c1 = (float)(-1.0 / 6.0); d.sinc_c2 = (float)(1.0 / 120.0);
    d.cos_c1 = -0.5f; d.cos_c2 = (float)(1.0 / 24.0); d.cos_c3 = (float)(-1.0 / 720.0);
    d.nsub = c->substeps * c->control_freq_inv;
    int blackout = 0;
    if (c->pomdp_mode == 1 || c->pomdp_mode == 3) {
        uint32_t r[4];
        draw(c->seed, GLOBAL_ENV, step, P_FLICKER, r);
        blackout = u01(r[0]) <= d.flicker_p;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t genv = (uint32_t)c->env_id_base + (uint32_t)i;
        float* rs = root + i * 13; float* T = thrust + i * 4; float* tg = target + i * 3; float* pr = params + i * 8;
        float p[3] = {rs[0], rs[1], rs[2]}, q[4] = {rs[3], rs[4], rs[5], rs[6]}, v[3] = {rs[7], rs[8], rs[9]}, w[3] = {rs[10], rs[11], rs[12]};
        int64_t prog = progress[i];
        const int rst = reset[i] != 0;
        uint32_t r4[4];
        int resample = rst || (!c->target_fixed && (prog % c->target_period) == 0);
        if (c->target_fixed) resample = 0;
        if (resample) {
            draw(c->seed, genv, step, P_TARGET, r4);
            for (int j = 0; j < 3; ++j) tg[j] = u01(r4[j]) * c->target_scale[j] + c->target_off[j];
        }
        if (rst) {
            draw(c->seed, genv, step, P_SPAWN, r4);
            for (int j = 0; j < 3; ++j) p[j] = c->spawn_base[j] + (c->spawn_range[j] * u01(r4[j]) + c->spawn_lo[j]);
            q[0] = q[1] = q[2] = 0.0f; q[3] = 1.0f;
            for (int j = 0; j < 3; ++j) { v[j] = 0.0f; w[j] = 0.0f; }
            prog = 0;
            if (c->fault_mode) {
                draw(c->seed, genv, step, P_FAULT, r4);
                fault[i * 2] = (int32_t)(r4[0] & 3u);
                fault[i * 2 + 1] = (int32_t)(((uint64_t)r4[1] * (uint32_t)c->max_episode_length) >> 32);
                pr[6] = c->fault_eff_lo + c->fault_eff_range * u01(r4[2]);
            }
            if (c->dr_enable) {
                uint32_t a[4], b[4], a2[4] = {0, 0, 0, 0}, b2[4] = {0, 0, 0, 0};
                int any_gauss = 0;
                for (int j = 0; j < OZL_DR_NUM; ++j) any_gauss |= c->dr[j].distribution == OZL_DR_GAUSSIAN;
                draw(c->seed, genv, step, P_DR0, a); draw(c->seed, genv, step, P_DR1, b);
                if (any_gauss) { draw(c->seed, genv, step, P_DR2, a2); draw(c->seed, genv, step, P_DR3, b2); }
                pr[0] = dr_apply(&c->dr[OZL_DR_MASS], c->mass, a[0], a2[0], step);
                pr[1] = dr_apply(&c->dr[OZL_DR_IXX], c->ixx, a[1], a2[1], step);
                pr[2] = dr_apply(&c->dr[OZL_DR_IYY], c->iyy, a[2], a2[2], step);
                pr[3] = dr_apply(&c->dr[OZL_DR_IZZ], c->izz, a[3], a2[3], step);
                pr[4] = dr_apply(&c->dr[OZL_DR_ARM], c->arm, b[0], b2[0], step);
                pr[5] = dr_apply(&c->dr[OZL_DR_THRUST_SCALE], 1.0f, b[1], b2[1], step);
                pr[7] = dr_apply(&c->dr[OZL_DR_YAW_KM], c->yaw_km, b[2], b2[2], step);
            }
        }
        /* thrust command (ouzelum.py:237-248) */
        float F[4];
        for (int k = 0; k < 4; ++k) {
            const float a = fminf(fmaxf(actions[i * 4 + k], -c->clip_actions), c->clip_actions);
            float t = T[k] + c->thrust_rate * a;
            t = fmaxf(fminf(t, c->thrust_max), 0.0f);
            F[k] = rst ? 0.0f : t;
            T[k] = F[k];
            F[k] = F[k] * pr[5];
        }
        const int fault_active = c->fault_mode && prog >= (int64_t)fault[i * 2 + 1];
        if (fault_active) F[fault[i * 2]] = F[fault[i * 2]] * pr[6];
        /* rigid body (body-frame angular velocity across the substeps, right-multiplied attitude update) */
        const float inv_m = 1.0f / pr[0];
        const float inertia[3] = {pr[1], pr[2], pr[3]};
        const float hi[3] = {d.h * (1.0f / pr[1]), d.h * (1.0f / pr[2]), d.h * (1.0f / pr[3])};
        const float fz = ((F[0] + F[1]) + F[2]) + F[3];
        const float tau_b[3] = {pr[4] * (((F[1] - F[0]) + F[2]) - F[3]), pr[4] * (((F[1] - F[0]) - F[2]) + F[3]),
                                pr[7] * (((F[2] - F[0]) - F[1]) + F[3])};
        float R[3][3], fw[3], tau_w[3], aw[3], rc[3], x[3], t3[3], wb[3], tb[3];
        quat_to_R(q, R);
        for (int j = 0; j < 3; ++j) fw[j] = R[j][2] * fz;
        mv(R, tau_b, tau_w);
        const float g[3] = {0.0f, 0.0f, c->gravity_z};
        const float kdm = c->lin_drag * inv_m;
        for (int j = 0; j < 3; ++j) { aw[j] = fmaf(fw[j], inv_m, g[j]); rc[j] = c->com_z * R[j][2]; x[j] = p[j] + rc[j]; }
        cross3(w, rc, t3);
        for (int j = 0; j < 3; ++j) { v[j] = v[j] + t3[j]; tb[j] = tau_b[j]; }
        mtv(R, w, wb);
        for (int s = 0; s < d.nsub; ++s) {
            for (int j = 0; j < 3; ++j) { v[j] = fmaf(d.h, fmaf(-kdm, v[j], aw[j]), v[j]); x[j] = fmaf(d.h, v[j], x[j]); }
            float iw[3], gy[3];
            for (int j = 0; j < 3; ++j) iw[j] = inertia[j] * wb[j];
            cross3(wb, iw, gy);
            for (int j = 0; j < 3; ++j) wb[j] = fmaf(hi[j], tb[j] - gy[j], wb[j]);
            float n2 = fmaf(wb[2], wb[2], fmaf(wb[1], wb[1], wb[0] * wb[0]));
            if (n2 > d.max_angvel2) {
                const float sc = c->max_angvel / sqrtf(n2);
                for (int j = 0; j < 3; ++j) wb[j] = wb[j] * sc;
                n2 = fmaf(wb[2], wb[2], fmaf(wb[1], wb[1], wb[0] * wb[0]));
            }
            const float th2 = d.hh2 * n2;
            const float sinc = fmaf(th2, fmaf(th2, d.sinc_c2, d.sinc_c1), 1.0f);
            const float cs = fmaf(th2, fmaf(th2, fmaf(th2, d.cos_c3, d.cos_c2), d.cos_c1), 1.0f);
            const float k = d.hh * sinc;
            const float dx = k * wb[0], dy = k * wb[1], dz = k * wb[2];
            const float nx = fmaf(q[3], dx, fmaf(cs, q[0], fmaf(q[1], dz, -(q[2] * dy))));
            const float ny = fmaf(q[3], dy, fmaf(cs, q[1], fmaf(q[2], dx, -(q[0] * dz))));
            const float nz = fmaf(q[3], dz, fmaf(cs, q[2], fmaf(q[0], dy, -(q[1] * dx))));
            const float nw = fmaf(q[3], cs, -fmaf(q[0], dx, fmaf(q[1], dy, q[2] * dz)));
            const float s2 = fmaf(nx, nx, fmaf(ny, ny, fmaf(nz, nz, nw * nw)));
            const float inv = fmaf(-0.5f, s2, 1.5f);
            q[0] = nx * inv; q[1] = ny * inv; q[2] = nz * inv; q[3] = nw * inv;
            quat_to_R(q, R);
            if (s + 1 < d.nsub) mtv(R, tau_w, tb);
        }
        mv(R, wb, w);
        for (int j = 0; j < 3; ++j) rc[j] = c->com_z * R[j][2];
        cross3(w, rc, t3);
        for (int j = 0; j < 3; ++j) { p[j] = x[j] - rc[j]; v[j] = v[j] - t3[j]; }
        /* post_physics_step: obs (ouzelum.py:280-285), reward (:302-332) */
        prog += 1;
        float o[13];
        const float dx = tg[0] - p[0], dy = tg[1] - p[1], dz = tg[2] - p[2];
        o[0] = dx * d.inv3; o[1] = dy * d.inv3; o[2] = dz * d.inv3;
        for (int j = 0; j < 4; ++j) o[3 + j] = q[j];
        for (int j = 0; j < 3; ++j) { o[7 + j] = v[j] * d.half; o[10 + j] = w[j] * d.inv_pi; }
        const float dist = sqrtf((dx * dx + dy * dy) + dz * dz);
        const float pos_r = 1.0f / (1.0f + dist * dist);
        const float ups_z = (2.0f * (q[3] * q[3]) - 1.0f) + (q[2] * q[2]) * 2.0f;
        const float tilt = fabsf(1.0f - ups_z);
        const float up_r = (1.0f / (1.0f + tilt * tilt)) * c->up_coef;
        const float spin = fabsf(w[2]);
        const float spin_r = 1.0f / (1.0f + spin * spin);
        const float reward = pos_r + pos_r * (up_r + spin_r);
        const int die = (dist > c->die_dist) || (p[2] < c->die_z);
        const int over = prog >= (int64_t)(c->max_episode_length - 1);
        const int rsout = over ? 1 : die;
        if (c->pomdp_mode != 0) {
            if (blackout) for (int j = 0; j < 13; ++j) o[j] = 0.0f;
            if (c->pomdp_mode >= 2)
                for (int k = 0; k < 4; ++k) {
                    draw(c->seed, genv, step, P_OBSNOISE + k, r4);
                    for (int j = 0; j < 4 && 4 * k + j < 13; ++j) o[4 * k + j] = o[4 * k + j] * (u01(r4[j]) * d.noise_range + d.noise_lo);
                }
        }
        for (int j = 0; j < 13; ++j) obs[i * 13 + j] = fminf(fmaxf(o[j], -c->clip_obs), c->clip_obs);
        rew[i] = reward;
        reset[i] = rsout;
        progress[i] = prog;
        if (timeout) timeout[i] = (uint8_t)(over && rsout);
        const float er = ep_ret[i] + reward;
        ep_ret[i] = rsout ? 0.0f : er;
        for (int j = 0; j < 3; ++j) { rs[j] = p[j]; rs[7 + j] = v[j]; rs[10 + j] = w[j]; }
        for (int j = 0; j < 4; ++j) rs[3 + j] = q[j];
    }
}

int ozl_oracle_cfg_size(void) { return (int)sizeof(ozl_cfg); }

#ifdef _OPENMP
#include <omp.h>
int ozl_oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
int ozl_oracle_set_threads(int n) { (void)n; return 1; }
#endif
