// K1q: the stock Quadcopter hover task (BASELINE config 1) as one fused kernel, one env per thread.
// CPU twin: oracle/quadcopter.py (QuadcopterOracle).  Reference (isaacgymenvs/tasks/quadcopter.py):
//   :280-299  reset_idx (root = (0,0,1) + U(-1.5,1.5),U(-1.5,1.5),U(-0.2,1.5); dof_pos = U(-0.2,0.2) x 8; dof_vel = 0)
//   :301-330  pre_physics_step (dof targets += dt*8*pi*a[0:8], clamp +-30 deg; thrusts += dt*200*a[8:12], clamp [0,2])
//   :359-370  compute_observations (21: (0,0,1)-pos)/3, quat, linvel/2, angvel/pi, 8 dof positions)
//   :386-418  compute_quadcopter_reward (up_reward 1/(1+tilt^2), die dist>3 or z<0.3, max_len 500)
//   :121-202  the procedurally built vehicle: chassis + 4 x (arm sphere, pitch hinge, rotor cylinder, roll hinge)
// The 9-body / 8-DOF PhysX articulation (PD joints, stiffness 1000) is replaced by ONE rigid body whose four thrust
// vectors tilt kinematically with the joint targets -- dynamics parity is unpinned/unattainable;
// observation / reward / reset arithmetic is exact.  All state is caller-owned AoS (this is the 256-env CPU-baseline config).
#include "internal.h"

namespace ozl {

struct QCfg {
    int64_t n;
    uint64_t seed, step;
    uint32_t env_id_base;
    int32_t max_episode_length, nsub;
    float clip_actions, clip_obs, dof_rate, dof_limit, thrust_rate, thrust_max, die_dist, die_z;
    float arm_r, rotor_off, cos_a[4], sin_a[4];
    DevCfg body;     // mass / inertia / integrator constants (only the rigid-body fields are used)
};

// sin / cos on [-0.6, 0.6] by fixed polynomials (|err| < 3e-11): identical in the oracle => bit-exact parity
__device__ __forceinline__ void sincos_small(float x, float& s, float& c) {
    const float x2 = x * x;
    s = x * (1.0f + x2 * (-1.6666667e-1f + x2 * (8.3333333e-3f + x2 * (-1.9841270e-4f + x2 * 2.7557319e-6f))));
    c = 1.0f + x2 * (-0.5f + x2 * (4.1666667e-2f + x2 * (-1.3888889e-3f + x2 * (2.4801587e-5f + x2 * -2.7557319e-7f))));
}

__device__ __forceinline__ void quadcopter_env(const QCfg& c, const uint64_t step, const int64_t i, const float* __restrict__ actions,
                                                    float* __restrict__ root13, float* __restrict__ dof_pos, float* __restrict__ dof_tgt,
                                                    float* __restrict__ thrust, float* __restrict__ obs, float* __restrict__ rew,
                                                    int64_t* __restrict__ reset, int64_t* __restrict__ progress,
                                                    uint8_t* __restrict__ timeout) {
    const uint32_t genv = c.env_id_base + (uint32_t)i;
    Env e;
    float* r = root13 + i * 13;
    for (int j = 0; j < 3; ++j) { e.p[j] = r[j]; e.v[j] = r[7 + j]; e.w[j] = r[10 + j]; }
    for (int j = 0; j < 4; ++j) { e.q[j] = r[3 + j]; e.T[j] = thrust[i * 4 + j]; }
    float dp[8], dt_[8];
    for (int j = 0; j < 8; ++j) { dp[j] = dof_pos[i * 8 + j]; dt_[j] = dof_tgt[i * 8 + j]; }
    int64_t prog = progress[i];
    const bool rst = reset[i] != 0;
    const DevCfg& b = c.body;
    e.mass = b.mass; e.ixx = b.ixx; e.iyy = b.iyy; e.izz = b.izz; e.arm = 0.f; e.ks = 1.f; e.km = 0.f;

    if (rst) {                                                                        // quadcopter.py:280-299
        const uint4 s0 = draw(c.seed, genv, step, P_SPAWN);
        e.p[0] = b.spawn_base[0] + (b.spawn_range[0] * u01(s0.x) + b.spawn_lo[0]);
        e.p[1] = b.spawn_base[1] + (b.spawn_range[1] * u01(s0.y) + b.spawn_lo[1]);
        e.p[2] = b.spawn_base[2] + (b.spawn_range[2] * u01(s0.z) + b.spawn_lo[2]);
        e.q[0] = e.q[1] = e.q[2] = 0.f; e.q[3] = 1.f;
        for (int j = 0; j < 3; ++j) { e.v[j] = 0.f; e.w[j] = 0.f; }
        const uint4 d0 = draw(c.seed, genv, step, P_QDOF0), d1 = draw(c.seed, genv, step, P_QDOF1);
        const uint32_t rr[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        for (int j = 0; j < 8; ++j) dp[j] = 0.4f * u01(rr[j]) + -0.2f;                // torch_rand_float(-0.2, 0.2)
        prog = 0;
    }
    // ---- pre_physics_step (quadcopter.py:301-330)
    float F[4];
    for (int j = 0; j < 8; ++j) {
        const float a = fminf(fmaxf(actions[i * 12 + j], -c.clip_actions), c.clip_actions);
        float t = dt_[j] + c.dof_rate * a;
        t = fmaxf(fminf(t, c.dof_limit), -c.dof_limit);
        dt_[j] = rst ? dp[j] : t;                                                     // :327 targets[reset] = dof_positions[reset]
    }
    for (int j = 0; j < 4; ++j) {
        const float a = fminf(fmaxf(actions[i * 12 + 8 + j], -c.clip_actions), c.clip_actions);
        float t = e.T[j] + c.thrust_rate * a;
        t = fmaxf(fminf(t, c.thrust_max), 0.0f);
        F[j] = rst ? 0.0f : t;
        e.T[j] = F[j];
    }
    // ---- wrench of four tilting rotors (kinematic joints: the rotor sits at its position target)
    float fb[3] = {0.f, 0.f, 0.f}, tau[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < 4; ++k) {
        float sp, cp, sr, cr;
        sincos_small(dt_[2 * k], sp, cp);
        sincos_small(dt_[2 * k + 1], sr, cr);
        const float dl[3] = {sp * cr, -sr, cp * cr};                                  // Ry(pitch) Rx(roll) e_z
        const float pl[3] = {c.arm_r + c.rotor_off * cp, 0.0f, -c.rotor_off * sp};    // arm + Ry(pitch) (rotor_off,0,0)
        const float d[3] = {c.cos_a[k] * dl[0] - c.sin_a[k] * dl[1], c.sin_a[k] * dl[0] + c.cos_a[k] * dl[1], dl[2]};
        const float p[3] = {c.cos_a[k] * pl[0] - c.sin_a[k] * pl[1], c.sin_a[k] * pl[0] + c.cos_a[k] * pl[1], pl[2]};
        const float f[3] = {F[k] * d[0], F[k] * d[1], F[k] * d[2]};
        float t3[3];
        cross3(p, f, t3);
        for (int j = 0; j < 3; ++j) { fb[j] = fb[j] + f[j]; tau[j] = tau[j] + t3[j]; }
    }
    simulate<false>(e, 0.0f, tau, b, fb);
    for (int j = 0; j < 8; ++j) dp[j] = dt_[j];
    // ---- post_physics_step: observations (quadcopter.py:359-370) and reward (:386-418)
    prog += 1;
    float o[21];
    o[0] = (0.0f - e.p[0]) * b.inv3; o[1] = (0.0f - e.p[1]) * b.inv3; o[2] = (1.0f - e.p[2]) * b.inv3;
    for (int j = 0; j < 4; ++j) o[3 + j] = e.q[j];
    for (int j = 0; j < 3; ++j) { o[7 + j] = e.v[j] * b.half; o[10 + j] = e.w[j] * b.inv_pi; }
    for (int j = 0; j < 8; ++j) o[13 + j] = dp[j];
    const float dz = 1.0f - e.p[2];
    const float dist = sqrtf((e.p[0] * e.p[0] + e.p[1] * e.p[1]) + dz * dz);
    const float pos_r = 1.0f / (1.0f + dist * dist);
    const float ups_z = (2.0f * (e.q[3] * e.q[3]) - 1.0f) + (e.q[2] * e.q[2]) * 2.0f;
    const float tilt = fabsf(1.0f - ups_z);
    const float up_r = 1.0f / (1.0f + tilt * tilt);
    const float spin = fabsf(e.w[2]);
    const float spin_r = 1.0f / (1.0f + spin * spin);
    const float reward = pos_r + pos_r * (up_r + spin_r);
    const bool die = (dist > c.die_dist) || (e.p[2] < c.die_z);
    const bool over = prog >= (int64_t)(c.max_episode_length - 1);
    const bool rs = over ? true : die;
    // ---- write back
    for (int j = 0; j < 3; ++j) { r[j] = e.p[j]; r[7 + j] = e.v[j]; r[10 + j] = e.w[j]; }
    for (int j = 0; j < 4; ++j) { r[3 + j] = e.q[j]; thrust[i * 4 + j] = e.T[j]; }
    for (int j = 0; j < 8; ++j) { dof_pos[i * 8 + j] = dp[j]; dof_tgt[i * 8 + j] = dt_[j]; }
    for (int j = 0; j < 21; ++j) obs[i * 21 + j] = fminf(fmaxf(o[j], -c.clip_obs), c.clip_obs);
    rew[i] = reward;
    reset[i] = rs ? 1 : 0;
    progress[i] = prog;
    if (timeout) timeout[i] = (over && rs) ? 1 : 0;
}

__global__ void __launch_bounds__(128)
quadcopter_step_kernel(const QCfg c, const float* __restrict__ actions, float* __restrict__ root13, float* __restrict__ dof_pos,
                       float* __restrict__ dof_tgt, float* __restrict__ thrust, float* __restrict__ obs, float* __restrict__ rew,
                       int64_t* __restrict__ reset, int64_t* __restrict__ progress, uint8_t* __restrict__ timeout,
                       unsigned long long* __restrict__ step_record, const uint32_t step_pad) {
    // step index: the host-passed value, or -- graph-capturable -- a device step-counter record (step_counter.cuh) read by one
    // thread per block and retired with one reduction per block at the end
    __shared__ uint64_t s_step;
    if (step_record && threadIdx.x == 0) s_step = read_step(step_record);
    __syncthreads();
    const uint64_t step = step_record ? s_step : c.step;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.n) quadcopter_env(c, step, i, actions, root13, dof_pos, dof_tgt, thrust, obs, rew, reset, progress, timeout);
    __syncthreads();
    if (step_record && threadIdx.x == 0) retire_units(step_record, 1ull + (blockIdx.x == 0 ? (unsigned long long)step_pad : 0ull));
}

}  // namespace ozl

using namespace ozl;

static unsigned record_shift(unsigned blocks) {      // 2^shift >= blocks: work units per step, padded to a power of two
    unsigned l = 0;
    while ((1ull << l) < blocks) ++l;
    return l;
}

extern "C" int ozl_quadcopter_step(const ozl_quadcopter_args* a, void* stream) {
    if (!a) return set_error("ozl_quadcopter_step: args is NULL");
    if (a->n <= 0) return set_error("ozl_quadcopter_step: n must be > 0");
    if (!a->actions12 || !a->root13 || !a->dof_pos8 || !a->dof_target8 || !a->thrust4 || !a->obs21 || !a->rew || !a->reset ||
        !a->progress)
        return set_error("ozl_quadcopter_step: NULL buffer");
    if (a->substeps <= 0 || !(a->mass > 1e-30f && a->mass < 1e30f) || !(a->ixx > 1e-30f && a->ixx < 1e30f) ||
        !(a->iyy > 1e-30f && a->iyy < 1e30f) || !(a->izz > 1e-30f && a->izz < 1e30f))
        return set_error("ozl_quadcopter_step: bad configuration");
    QCfg c;
    memset(&c, 0, sizeof(c));
    c.n = a->n; c.seed = a->seed; c.step = a->step; c.env_id_base = (uint32_t)a->env_id_base;
    c.max_episode_length = a->max_episode_length; c.nsub = a->substeps;
    c.clip_actions = a->clip_actions; c.clip_obs = a->clip_obs;
    c.dof_rate = (float)((double)a->dt * 8.0 * M_PI);            // quadcopter.py:310-311  dt * dof_action_speed_scale
    c.dof_limit = (float)(30.0 * M_PI / 180.0);                  // hinge range -30..30 degrees, quadcopter.py:180,199
    c.thrust_rate = (float)((double)a->dt * 200.0);              // quadcopter.py:314-315
    c.thrust_max = 2.0f;                                         // quadcopter.py:88
    c.die_dist = 3.0f; c.die_z = 0.3f;                           // quadcopter.py:412-413
    c.arm_r = (float)(0.1 + 0.25 * 0.01);                        // chassis_radius + 0.25 rotor_arm_radius, quadcopter.py:156
    c.rotor_off = (float)(0.04 + 0.25 * 0.01);                   // rotor_radius + 0.25 rotor_arm_radius,  quadcopter.py:159
    for (int k = 0; k < 4; ++k) {
        const double ang = (0.25 + 0.5 * k) * M_PI;              // quadcopter.py:161
        c.cos_a[k] = (float)cos(ang); c.sin_a[k] = (float)sin(ang);
    }
    DevCfg& b = c.body;
    b.nsub = a->substeps;
    b.mass = a->mass; b.ixx = a->ixx; b.iyy = a->iyy; b.izz = a->izz; b.com_z = 0.0f;
    b.max_angvel = (float)(4.0 * M_PI);
    b.max_angvel2 = (float)((double)b.max_angvel * (double)b.max_angvel);
    b.gravity_z = a->gravity_z; b.lin_drag = 0.0f;
    const double h = (double)a->dt / (double)a->substeps;
    b.h = (float)h; b.hh = (float)(0.5 * h); b.hh2 = b.hh * b.hh;
    b.sinc_c1 = (float)(-1.0 / 6.0); b.sinc_c2 = (float)(1.0 / 120.0);
    b.cos_c1 = -0.5f; b.cos_c2 = (float)(1.0 / 24.0); b.cos_c3 = (float)(-1.0 / 720.0);
    b.inv3 = 1.0f / 3.0f; b.half = 0.5f; b.inv_pi = 1.0f / (float)M_PI;
    const float sb[3] = {0.f, 0.f, 1.f}, sl[3] = {-1.5f, -1.5f, -0.2f}, sr[3] = {3.0f, 3.0f, (float)(1.5 - (-0.2))};
    for (int j = 0; j < 3; ++j) { b.spawn_base[j] = sb[j]; b.spawn_lo[j] = sl[j]; b.spawn_range[j] = sr[j]; }
    const unsigned blocks = (unsigned)((a->n + 127) / 128);
    if (a->step_record && ((uintptr_t)a->step_record & 15)) return set_error("ozl_quadcopter_step: step_record must be 16-byte aligned");
    quadcopter_step_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(
        c, a->actions12, a->root13, a->dof_pos8, a->dof_target8, a->thrust4, a->obs21, a->rew, a->reset, a->progress, a->timeout,
        (unsigned long long*)a->step_record, (uint32_t)((1ull << record_shift(blocks)) - blocks));
    return check_cuda(cudaGetLastError(), "quadcopter_step_kernel");
}

// Stand-alone step-counter record (step_counter.cuh) for kernels that have no env handle (ozl_quadcopter_step): 16 bytes of
// device memory owned by the caller; `n_envs` fixes the number of 128-env blocks per step.
extern "C" int ozl_step_record_init(uint64_t* record_dev, int64_t n_envs, uint64_t step, void* stream) {
    if (!record_dev) return set_error("ozl_step_record_init: record is NULL");
    if (n_envs <= 0) return set_error("ozl_step_record_init: n_envs must be > 0");
    if ((uintptr_t)record_dev & 15) return set_error("ozl_step_record_init: record must be 16-byte aligned");
    if (step > kStepBaseMask) return set_error("ozl_step_record_init: step out of range");
    const unsigned long long w[2] = {step_word0((unsigned long long)step, record_shift((unsigned)((n_envs + 127) / 128))), 0ull};
    return check_cuda(cudaMemcpyAsync(record_dev, w, sizeof(w), cudaMemcpyHostToDevice, (cudaStream_t)stream), "cudaMemcpyAsync");
}

extern "C" int ozl_step_record_read(const uint64_t* record_dev, uint64_t* out, void* stream) {
    if (!record_dev || !out) return set_error("ozl_step_record_read: NULL argument");
    unsigned long long w[2] = {0, 0};
    if (check_cuda(cudaMemcpyAsync(w, record_dev, sizeof(w), cudaMemcpyDeviceToHost, (cudaStream_t)stream), "cudaMemcpyAsync")) return 1;
    if (check_cuda(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize")) return 1;
    *out = step_from_words(w);
    return 0;
}
