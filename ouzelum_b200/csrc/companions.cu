// Companion kernels of the env step: Lee controllers (K4), PV Kalman filter (K3), attitude EKF (K2), sensor-fault
// model on arbitrary vectors (K7), episode statistics (part of K6).  One env per thread; see the .cuh files for the math.
#include "internal.h"
#include "lee_control.cuh"
#include "filters.cuh"
#include "glue.cuh"

namespace ozl {

static inline unsigned nblk(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }

// ------------------------------------------------------------------------------------------------ K4
__global__ void __launch_bounds__(128)
lee_control_kernel(int mode, int64_t n, const float* __restrict__ state13, const float* __restrict__ cmd4, const LeeGains g,
                   float* __restrict__ thrust, float* __restrict__ torque3, float4* __restrict__ wrench4, float thrust_scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* s = state13 + i * 13;
    const float p[3] = {s[0], s[1], s[2]}, q[4] = {s[3], s[4], s[5], s[6]}, v[3] = {s[7], s[8], s[9]}, w[3] = {s[10], s[11], s[12]};
    const float4 c4 = reinterpret_cast<const float4*>(cmd4)[i];
    const float cmd[4] = {c4.x * g.scale[0], c4.y * g.scale[1], c4.z * g.scale[2], c4.w * g.scale[3]};   // controller.py:47
    float th, tq[3];
    lee_control(mode, p, q, v, w, cmd, g, th, tq);
    if (thrust) thrust[i] = th;
    if (torque3) { torque3[i * 3 + 0] = tq[0]; torque3[i * 3 + 1] = tq[1]; torque3[i * 3 + 2] = tq[2]; }
    // forces[:,0,2] = mg * thrust ; torques[:,0] = torque   (lee_landed.py:313-314, ekf_lee_landed.py:504-505)
    if (wrench4) wrench4[i] = make_float4(thrust_scale * th, tq[0], tq[1], tq[2]);
}

// ------------------------------------------------------------------------------------------------ K3
struct PVArgs {
    int64_t n;
    float* x;            // [9][n]
    float* P;            // [81][n]
    const float* accel;  // [n,3]
    const float* quat;   // [n,4]
    const float* pos_meas;   // [n,3] or null
    const float* vel_meas;   // [n,3] or null
    const uint8_t* pos_mask; // [n] or null
    const uint8_t* vel_mask; // [n] or null
    float dt, dt2;
    float acc_var[3], pos_var[3], vel_var[3];
    int flip_qw, do_predict;
    uint32_t pos_period, pos_phase, vel_period, vel_phase;   // used when the mask pointer is null (0 = never)
    uint64_t iter_base;                                        // global iteration index of env 0 (tasks/ekf_lee_landed.py:425-440)
};

// Stand-alone PV filter step: the block's [81][128] covariance slice is staged in shared memory (coalesced plane loads) and the
// filter runs on it with the same device functions as the fused EKFLeeLanded kernel (filters.cuh) -- identical bits.
constexpr int kPvBlock = 128;
__global__ void __launch_bounds__(kPvBlock)
pv_step_kernel(const PVArgs a) {
    __shared__ float s_P[81 * kPvBlock];
    const int64_t i = (int64_t)blockIdx.x * kPvBlock + threadIdx.x;
    if (i >= a.n) return;                              // no block-level synchronisation below: every thread works on its own column
    PVShared<kPvBlock> s;
    s.P = s_P + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 9; ++k) s.x[k] = a.x[(int64_t)k * a.n + i];
#pragma unroll 9
    for (int k = 0; k < 81; ++k) s_P[k * kPvBlock + threadIdx.x] = a.P[(int64_t)k * a.n + i];
    if (a.do_predict) {
        const float acc[3] = {a.accel[i * 3], a.accel[i * 3 + 1], a.accel[i * 3 + 2]};
        const float4 q4 = reinterpret_cast<const float4*>(a.quat)[i];
        float q[4];
        if (a.flip_qw) { q[0] = q4.w; q[1] = q4.x; q[2] = q4.y; q[3] = q4.z; }     // xyzw -> wxyz (PVFilter.py:32-33)
        else { q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; q[3] = q4.w; }
        pv_predict(s, acc, q, a.dt, a.dt2, a.acc_var);
    }
    const uint64_t k = a.iter_base + (uint64_t)i;
    bool pos_fix = false, vel_fix = false;
    if (a.pos_meas) pos_fix = a.pos_mask ? (a.pos_mask[i] != 0) : (a.pos_period && (k % a.pos_period) == a.pos_phase);
    if (a.vel_meas) vel_fix = a.vel_mask ? (a.vel_mask[i] != 0) : (a.vel_period && (k % a.vel_period) == a.vel_phase);
    if (pos_fix) {                                                                  // ekf_lee_landed.py:428-433
        const float z[3] = {a.pos_meas[i * 3], a.pos_meas[i * 3 + 1], a.pos_meas[i * 3 + 2]};
        pv_correct<0>(s, z, a.pos_var);
    }
    if (vel_fix) {                                                                  // ekf_lee_landed.py:435-440
        const float z[3] = {a.vel_meas[i * 3], a.vel_meas[i * 3 + 1], a.vel_meas[i * 3 + 2]};
        pv_correct<3>(s, z, a.vel_var);
    }
#pragma unroll
    for (int kk = 0; kk < 9; ++kk) a.x[(int64_t)kk * a.n + i] = s.x[kk];
#pragma unroll 9
    for (int kk = 0; kk < 81; ++kk) a.P[(int64_t)kk * a.n + i] = s_P[kk * kPvBlock + threadIdx.x];
}

__global__ void pv_init_kernel(int64_t n, float* x, float* P) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < 9; ++k) x[(int64_t)k * n + i] = 0.0f;                        // PVFilter.py:11
    for (int k = 0; k < 81; ++k) P[(int64_t)k * n + i] = (k / 9 == k % 9) ? 1000.0f : 0.0f;   // PVFilter.py:12
}

// re-seed the state of reset envs with the true position / velocity (tasks/ekf_lee_landed.py:353-358)
__global__ void pv_reset_kernel(int64_t n, float* x, const int64_t* flags, const float* root13) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (flags && flags[i] == 0)) return;
    const float* r = root13 + i * 13;
    for (int k = 0; k < 3; ++k) {
        x[(int64_t)k * n + i] = r[k];
        x[(int64_t)(3 + k) * n + i] = r[7 + k];
        x[(int64_t)(6 + k) * n + i] = 0.0f;
    }
}

// ------------------------------------------------------------------------------------------------ K2
__global__ void __launch_bounds__(128)
ekf_update_kernel(int64_t n, double* __restrict__ q, double* __restrict__ P, const float* __restrict__ gyr,
                  const float* __restrict__ ang, int ang_xyzw, double Dt, double g_noise, double s_eps,
                  float4* __restrict__ q_f32_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EKF4 s;
#pragma unroll
    for (int k = 0; k < 4; ++k) s.q[k] = q[(int64_t)k * n + i];
#pragma unroll
    for (int k = 0; k < 16; ++k) s.P[k / 4][k % 4] = P[(int64_t)k * n + i];
    const double g[3] = {(double)gyr[i * 3], (double)gyr[i * 3 + 1], (double)gyr[i * 3 + 2]};
    const float4 a4 = reinterpret_cast<const float4*>(ang)[i];
    double a[4];
    if (ang_xyzw) { a[0] = a4.w; a[1] = a4.x; a[2] = a4.y; a[3] = a4.z; }        // root_quats[idx,[3,0,1,2]]
    else { a[0] = a4.x; a[1] = a4.y; a[2] = a4.z; a[3] = a4.w; }
    ekf_update(s, g, a, Dt, g_noise, s_eps);
#pragma unroll
    for (int k = 0; k < 4; ++k) q[(int64_t)k * n + i] = s.q[k];
#pragma unroll
    for (int k = 0; k < 16; ++k) P[(int64_t)k * n + i] = s.P[k / 4][k % 4];
    // torch.Tensor(self.Q_state): float64 -> float32 wxyz, the PV filter's orientation input (ekf_lee_landed.py:402)
    if (q_f32_out) q_f32_out[i] = make_float4((float)s.q[0], (float)s.q[1], (float)s.q[2], (float)s.q[3]);
}

__global__ void ekf_init_kernel(int64_t n, double* q, double* P) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < 4; ++k) q[(int64_t)k * n + i] = (k == 0) ? 1.0 : 0.0;
    for (int k = 0; k < 16; ++k) P[(int64_t)k * n + i] = (k / 4 == k % 4) ? 1.0 : 0.0;      // ahrs_ekf.py:995
}

// Q_state[ids] = root_quats[ids][:, [3,0,1,2]]  (tasks/ekf_lee_landed.py:349-352)
__global__ void ekf_set_q_kernel(int64_t n, double* q, const float* quat_xyzw, const int64_t* flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (flags && flags[i] == 0)) return;
    const float4 v = reinterpret_cast<const float4*>(quat_xyzw)[i];
    q[i] = v.w; q[n + i] = v.x; q[2 * n + i] = v.y; q[3 * n + i] = v.z;
}

// ------------------------------------------------------------------------------------------------ K7 (stand-alone)
// POMDPWrapper.observation on an [n, d] float32 array (utils/POMDP.py:23-42): flicker = one draw per call blacks out
// EVERY env; noise = element-wise U(1-s, 1+s).  Draws: philox(seed, env, step, P_OBSNOISE + j/4)[j%4]; the flicker draw
// uses env word GLOBAL_ENV.  `stream_id` separates several uses within one step (gyro / accel / pos / vel ...).
__global__ void pomdp_kernel(int64_t n, int d, int mode, float flicker_p, float noise_lo, float noise_range, uint64_t seed,
                             uint64_t step, uint32_t env_id_base, uint32_t stream_id, const float* __restrict__ in,
                             float* __restrict__ out, const unsigned long long* __restrict__ step_ptr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (step_ptr) step = read_step(step_ptr);   // follow a handle's device step counter (3-word record, step_counter.cuh)
    bool blackout = false;
    if (mode == 1 || mode == 3) {
        const uint4 r = draw(seed, GLOBAL_ENV, step, P_FLICKER + (stream_id << 8));
        blackout = u01(r.x) <= flicker_p;
    }
    const uint32_t genv = env_id_base + (uint32_t)i;
    for (int j0 = 0; j0 < d; j0 += 4) {
        uint4 r = make_uint4(0, 0, 0, 0);
        if (mode >= 2) r = draw(seed, genv, step, (P_OBSNOISE + (uint32_t)(j0 >> 2)) + (stream_id << 8));
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
        for (int j = 0; j < 4 && j0 + j < d; ++j) {
            float v = blackout ? 0.0f : in[i * d + j0 + j];
            if (mode >= 2) v = __fmul_rn(v, __fadd_rn(__fmul_rn(u01(rr[j]), noise_range), noise_lo));   // no FMA: bit-exact vs oracle
            out[i * d + j0 + j] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------ DR noise lambdas
// The observation / action noise of the reference's domain randomisation (tasks/base/vec_task.py:576-646): per element
//     gaussian:  out = op(x, (corr * var_corr + mu_corr) + randn * var + mu)
//     uniform:   out = op(x, (corr * (hi_corr - lo_corr) + lo_corr) + rand * (hi - lo) + lo)
// with `corr` a standard-normal sample drawn ONCE per randomisation event and kept until the next one (params['corr'], :612-616,
// :637-641) and randn / rand fresh every step; op = + or *.  The schedule (:583-609, :621-635) is applied by the host when the
// event happens, so a / b / a_corr / b_corr arrive already scaled.  Draws: philox(seed, env, step, P_LAMBDA + 16 which + j/4) for the
// fresh sample and philox(seed, env, corr_epoch, P_LAMBDA_CORR + 16 which + j/4) for the correlated one (stateless: the same
// value every step of the event); normals by Box-Muller in float64 from the 24-bit uniforms (u1 in (0,1]), rounded to float32 once.
// `clip` > 0 clamps the result (the observation clamp of vec_task.py:353 comes after the noise).
constexpr uint32_t P_LAMBDA = 32, P_LAMBDA_CORR = 64;
__device__ __forceinline__ void box_muller4(const uint4 r, float z[4]) {
    const double k = 5.9604644775390625e-08, two_pi = 6.283185307179586;
    const double u1 = ((double)(r.x >> 8) + 1.0) * k, u2 = (double)(r.y >> 8) * k;
    const double u3 = ((double)(r.z >> 8) + 1.0) * k, u4 = (double)(r.w >> 8) * k;
    const double ra = sqrt(-2.0 * log(u1)), rb = sqrt(-2.0 * log(u3));
    double sa, ca, sb, cb;
    sincos(two_pi * u2, &sa, &ca);
    sincos(two_pi * u4, &sb, &cb);
    z[0] = (float)(ra * ca); z[1] = (float)(ra * sa); z[2] = (float)(rb * cb); z[3] = (float)(rb * sb);
}
struct NoiseLambda { int dist, op; float a, b, a_corr, b_corr; };
__global__ void noise_lambda_kernel(int64_t n, int width, float* __restrict__ x, const NoiseLambda s, float clip, uint64_t seed,
                                    uint64_t step, const unsigned long long* __restrict__ step_ptr, long long step_offset,
                                    uint64_t corr_epoch, uint32_t env_id_base, uint32_t which) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (step_ptr) step = (uint64_t)((long long)read_step(step_ptr) + step_offset);
    const uint32_t genv = env_id_base + (uint32_t)i;
    const bool gauss = s.dist == OZL_DR_GAUSSIAN;
    const float corr_scale = gauss ? s.b_corr : (s.b_corr - s.a_corr), scale = gauss ? s.b : (s.b - s.a);
    for (int j0 = 0; j0 < width; j0 += 4) {
        const uint32_t g = (uint32_t)(j0 >> 2) + 16u * which;
        const uint4 r = draw(seed, genv, step, P_LAMBDA + g), rc = draw(seed, genv, corr_epoch, P_LAMBDA_CORR + g);
        float zc[4], zf[4];
        box_muller4(rc, zc);
        if (gauss) box_muller4(r, zf);
        else { zf[0] = u01(r.x); zf[1] = u01(r.y); zf[2] = u01(r.z); zf[3] = u01(r.w); }
        for (int j = 0; j < 4 && j0 + j < width; ++j) {
            const float corr = zc[j] * corr_scale + s.a_corr;
            const float noise = (corr + zf[j] * scale) + s.a;
            float v = x[i * width + j0 + j];
            v = s.op == OZL_DR_ADDITIVE ? v + noise : v * noise;
            if (clip > 0.0f) v = fminf(fmaxf(v, -clip), clip);
            x[i * width + j0 + j] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------ E3 glue
// Sensor front-end of EKFLeeLanded.pre_physics_step (tasks/ekf_lee_landed.py:345-346,366-375,397-406): from the true root
// state builds, per env, sensors[16] = accel(3) | gyr(3) | ang xyzw(4) | pos(3) | vel(3), optionally through the
// sensor-fault model (after the warm-up).  accel = (linvel - prev_linvel)/dt with +9.8 on z (the reference adds 9.8, not
// 9.81, in place -- :367 -- so the PV filter receives the gravity-added, un-rotated vector).  prev_linvel is updated (:454).
struct FrontArgs {
    int64_t n;
    const float* root13;
    float* prev_linvel;      // [n,3] in/out
    float* sensors;          // [n,16] out
    float dt, inv_dt;        // inv_dt = 1.0f / dt: torch-CUDA evaluates `tensor / python_scalar` as a multiplication by the float32 reciprocal
    FaultCfg f;
    uint32_t env_id_base;
};
__global__ void sensor_frontend_kernel(const FrontArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const float* r = a.root13 + i * 13;
    const uint32_t genv = a.env_id_base + (uint32_t)i;
    float acc[3], gyr[3], ang[4], pos[3], vel[3];
    for (int j = 0; j < 3; ++j) {
        acc[j] = (r[7 + j] - a.prev_linvel[i * 3 + j]) * a.inv_dt;       // :345-346 (dv / dt on a CUDA tensor)
        gyr[j] = r[10 + j]; pos[j] = r[j]; vel[j] = r[7 + j];
    }
    acc[2] = acc[2] + 9.8f;                                              // :367
    for (int j = 0; j < 4; ++j) ang[j] = r[3 + j];
    sensor_fault(a.f, genv, 1, false, gyr, 3);                                   // :374
    sensor_fault(a.f, genv, 3, true, ang, 4);                                    // :383
    sensor_fault(a.f, genv, 4, false, acc, 3);                                   // :401
    sensor_fault(a.f, genv, 5, false, pos, 3);                                   // :403
    sensor_fault(a.f, genv, 6, false, vel, 3);                                   // :404
    float* o = a.sensors + i * 16;
    for (int j = 0; j < 3; ++j) { o[j] = acc[j]; o[3 + j] = gyr[j]; o[10 + j] = pos[j]; o[13 + j] = vel[j]; }
    for (int j = 0; j < 4; ++j) o[6 + j] = ang[j];
    for (int j = 0; j < 3; ++j) a.prev_linvel[i * 3 + j] = r[7 + j];     // :454
}

// Waypoint ("carrot") logic + controller input assembly (tasks/ekf_lee_landed.py:458-503).
//   warm-up: waypoint = target, controller runs on the true state
//   after:   if waypoint_dist < 0.5 or > 1.0: waypoint = pos + 0.75 * unit(target + (0,0,0.7) - pos)
//            if target_dist < 0.75:           waypoint = target + (0,0,0.09)
//            controller state = [PV position, true quat, PV velocity, true angvel]
// The reference guards the carrot update with a GLOBAL `(waypoint_dist == 0).any()` (:474); here the guard is per env.
__global__ void waypoint_kernel(int64_t n, const float* __restrict__ root13, const float* __restrict__ pv_x,
                                const float* __restrict__ target3, float* __restrict__ waypoint3, int warm,
                                float* __restrict__ est13, float4* __restrict__ cmd4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = root13 + i * 13;
    const float p[3] = {r[0], r[1], r[2]}, t[3] = {target3[i * 3], target3[i * 3 + 1], target3[i * 3 + 2]};
    float w[3] = {waypoint3[i * 3], waypoint3[i * 3 + 1], waypoint3[i * 3 + 2]};
    waypoint_update(p, t, w, warm != 0);
    for (int j = 0; j < 3; ++j) waypoint3[i * 3 + j] = w[j];
    cmd4[i] = make_float4(w[0], w[1], w[2], 0.0f);                                         // :488
    float* e = est13 + i * 13;
    for (int j = 0; j < 13; ++j) e[j] = r[j];
    if (!warm) {                                                                            // :493-497
        for (int j = 0; j < 3; ++j) { e[j] = pv_x[(int64_t)j * n + i]; e[7 + j] = pv_x[(int64_t)(3 + j) * n + i]; }
    }
}

// ------------------------------------------------------------------------------------------------ episode statistics
// RecordEpisodeStatisticsTorch.step (RPO-LSTM/utils.py:20-35): 6 element-wise launches -> 1
__global__ void episode_stats_kernel(int64_t n, const float* __restrict__ rew, const int64_t* __restrict__ done,
                                     float* __restrict__ ep_ret, int32_t* __restrict__ ep_len, float* __restrict__ ret_out,
                                     int32_t* __restrict__ len_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float r = ep_ret[i] + rew[i];
    const int32_t l = ep_len[i] + 1;
    ret_out[i] = r;
    len_out[i] = l;
    const int64_t dn = done[i];
    ep_ret[i] = r * (float)(1 - dn);          // episode_returns *= 1 - dones
    ep_len[i] = l * (int32_t)(1 - dn);
}

}  // namespace ozl

using namespace ozl;

#define OZL_N_CHECK(name)                                        \
    if (n <= 0) return set_error(name ": n must be > 0");        \
    cudaStream_t st = (cudaStream_t)stream;

extern "C" int ozl_lee_control(int32_t mode, int64_t n, const float* state13, const float* cmd4, const float* gains16,
                               float* thrust, float* torque3, void* stream) {
    OZL_N_CHECK("ozl_lee_control");
    if (mode < 0 || mode > 2) return set_error("Invalid controller name: mode %d (0 position, 1 velocity, 2 attitude)", mode);
    if (!state13 || !cmd4 || !gains16 || !thrust || !torque3) return set_error("ozl_lee_control: NULL buffer");
    if ((uintptr_t)cmd4 & 15) return set_error("ozl_lee_control: cmd4 must be 16-byte aligned");
    LeeGains g;
    for (int k = 0; k < 3; ++k) { g.kP[k] = gains16[k]; g.kV[k] = gains16[3 + k]; g.kR[k] = gains16[6 + k]; g.kO[k] = gains16[9 + k]; }
    for (int k = 0; k < 4; ++k) g.scale[k] = gains16[12 + k];
    lee_control_kernel<<<nblk(n, 128), 128, 0, st>>>(mode, n, state13, cmd4, g, thrust, torque3, nullptr, 0.0f);
    return check_cuda(cudaGetLastError(), "lee_control_kernel");
}

extern "C" int ozl_lee_wrench(int32_t mode, int64_t n, const float* state13, const float* cmd4, const float* gains16,
                              float thrust_scale, float* wrench4, void* stream) {
    OZL_N_CHECK("ozl_lee_wrench");
    if (mode < 0 || mode > 2) return set_error("Invalid controller name: mode %d (0 position, 1 velocity, 2 attitude)", mode);
    if (!state13 || !cmd4 || !gains16 || !wrench4) return set_error("ozl_lee_wrench: NULL buffer");
    if (((uintptr_t)cmd4 & 15) || ((uintptr_t)wrench4 & 15)) return set_error("ozl_lee_wrench: cmd4/wrench4 must be 16-byte aligned");
    LeeGains g;
    for (int k = 0; k < 3; ++k) { g.kP[k] = gains16[k]; g.kV[k] = gains16[3 + k]; g.kR[k] = gains16[6 + k]; g.kO[k] = gains16[9 + k]; }
    for (int k = 0; k < 4; ++k) g.scale[k] = gains16[12 + k];
    lee_control_kernel<<<nblk(n, 128), 128, 0, st>>>(mode, n, state13, cmd4, g, nullptr, nullptr, (float4*)wrench4, thrust_scale);
    return check_cuda(cudaGetLastError(), "lee_control_kernel");
}

extern "C" int ozl_pv_init(int64_t n, float* x9xN, float* P81xN, void* stream) {
    OZL_N_CHECK("ozl_pv_init");
    if (!x9xN || !P81xN) return set_error("ozl_pv_init: NULL buffer");
    pv_init_kernel<<<nblk(n, 256), 256, 0, st>>>(n, x9xN, P81xN);
    return check_cuda(cudaGetLastError(), "pv_init_kernel");
}

extern "C" int ozl_pv_reset(int64_t n, float* x9xN, const int64_t* flags, const float* root13, void* stream) {
    OZL_N_CHECK("ozl_pv_reset");
    if (!x9xN || !root13) return set_error("ozl_pv_reset: NULL buffer");
    pv_reset_kernel<<<nblk(n, 256), 256, 0, st>>>(n, x9xN, flags, root13);
    return check_cuda(cudaGetLastError(), "pv_reset_kernel");
}

extern "C" int ozl_pv_step(const ozl_pv_args* a, void* stream) {
    if (!a) return set_error("ozl_pv_step: args is NULL");
    const int64_t n = a->n;
    OZL_N_CHECK("ozl_pv_step");
    if (!a->x9xN || !a->P81xN) return set_error("ozl_pv_step: NULL state buffer");
    if (a->do_predict && (!a->accel3 || !a->quat4)) return set_error("ozl_pv_step: predict needs accel3 and quat4");
    if (a->quat4 && ((uintptr_t)a->quat4 & 15)) return set_error("ozl_pv_step: quat4 must be 16-byte aligned");
    PVArgs k;
    k.n = n; k.x = a->x9xN; k.P = a->P81xN; k.accel = a->accel3; k.quat = a->quat4;
    k.pos_meas = a->pos_meas3; k.vel_meas = a->vel_meas3; k.pos_mask = a->pos_mask; k.vel_mask = a->vel_mask;
    k.dt = a->dt;
    k.dt2 = (float)((double)a->dt * (double)a->dt);       // python `dt**2` (double), then float32 against the tensor
    for (int j = 0; j < 3; ++j) {
        k.acc_var[j] = a->acc_var[j];
        k.pos_var[j] = a->pos_var_given ? a->pos_var[j] : 0.0f;       // PVFilter.py:96-99
        k.vel_var[j] = a->vel_var_given ? a->vel_var[j] : 0.0f;       // PVFilter.py:76-79 (the reference passes gps_var=None => 0)
    }
    k.flip_qw = a->flip_qw; k.do_predict = a->do_predict;
    k.pos_period = a->pos_period; k.pos_phase = a->pos_phase; k.vel_period = a->vel_period; k.vel_phase = a->vel_phase;
    k.iter_base = a->iter_base;
    pv_step_kernel<<<nblk(n, 128), 128, 0, st>>>(k);
    return check_cuda(cudaGetLastError(), "pv_step_kernel");
}

extern "C" int ozl_ekf_init(int64_t n, double* q4xN, double* P16xN, void* stream) {
    OZL_N_CHECK("ozl_ekf_init");
    if (!q4xN || !P16xN) return set_error("ozl_ekf_init: NULL buffer");
    ekf_init_kernel<<<nblk(n, 256), 256, 0, st>>>(n, q4xN, P16xN);
    return check_cuda(cudaGetLastError(), "ekf_init_kernel");
}

extern "C" int ozl_ekf_set_q(int64_t n, double* q4xN, const float* quat_xyzw, const int64_t* flags, void* stream) {
    OZL_N_CHECK("ozl_ekf_set_q");
    if (!q4xN || !quat_xyzw) return set_error("ozl_ekf_set_q: NULL buffer");
    ekf_set_q_kernel<<<nblk(n, 256), 256, 0, st>>>(n, q4xN, quat_xyzw, flags);
    return check_cuda(cudaGetLastError(), "ekf_set_q_kernel");
}

extern "C" int ozl_ekf_update(int64_t n, double* q4xN, double* P16xN, const float* gyr3, const float* ang4, int32_t ang_xyzw,
                              double Dt, double g_noise, float* q_wxyz_f32_out, void* stream) {
    OZL_N_CHECK("ozl_ekf_update");
    if (!q4xN || !P16xN || !gyr3 || !ang4) return set_error("ozl_ekf_update: NULL buffer");
    if ((uintptr_t)ang4 & 15) return set_error("ozl_ekf_update: ang4 must be 16-byte aligned");
    if ((uintptr_t)q_wxyz_f32_out & 15) return set_error("ozl_ekf_update: q_wxyz_f32_out must be 16-byte aligned");
    ekf_update_kernel<<<nblk(n, 128), 128, 0, st>>>(n, q4xN, P16xN, gyr3, ang4, ang_xyzw, Dt, g_noise, 0.0000001,
                                                    (float4*)q_wxyz_f32_out);
    return check_cuda(cudaGetLastError(), "ekf_update_kernel");
}

extern "C" int ozl_sensor_frontend(int64_t n, const float* root13, float* prev_linvel3, float* sensors16, float dt,
                                   int32_t mode, float pomdp_prob, uint64_t seed, uint64_t step, int64_t env_id_base,
                                   void* stream) {
    OZL_N_CHECK("ozl_sensor_frontend");
    if (!root13 || !prev_linvel3 || !sensors16) return set_error("ozl_sensor_frontend: NULL buffer");
    if (mode < 0 || mode > 3) return set_error("pomdp was not in ['flicker', 'random_noise', 'flickering_and_random_noise']!");
    FrontArgs a;
    a.n = n; a.root13 = root13; a.prev_linvel = prev_linvel3; a.sensors = sensors16; a.dt = dt; a.inv_dt = 1.0f / dt; a.f.mode = mode;
    a.f.flicker_p = (mode == 3) ? 0.1f : pomdp_prob;
    const float lo = (float)(1.0 - (double)pomdp_prob), hi = (float)(1.0 + (double)pomdp_prob);
    a.f.noise_lo = lo; a.f.noise_range = hi - lo;
    a.f.seed = seed; a.f.step = step; a.env_id_base = (uint32_t)env_id_base;
    sensor_frontend_kernel<<<nblk(n, 256), 256, 0, st>>>(a);
    return check_cuda(cudaGetLastError(), "sensor_frontend_kernel");
}

extern "C" int ozl_waypoint_command(int64_t n, const float* root13, const float* pv_x9xN, const float* target3,
                                    float* waypoint3, int32_t warmup, float* est13, float* cmd4, void* stream) {
    OZL_N_CHECK("ozl_waypoint_command");
    if (!root13 || !target3 || !waypoint3 || !est13 || !cmd4 || (!warmup && !pv_x9xN)) return set_error("ozl_waypoint_command: NULL buffer");
    if ((uintptr_t)cmd4 & 15) return set_error("ozl_waypoint_command: cmd4 must be 16-byte aligned");
    waypoint_kernel<<<nblk(n, 256), 256, 0, st>>>(n, root13, pv_x9xN, target3, waypoint3, warmup, est13, (float4*)cmd4);
    return check_cuda(cudaGetLastError(), "waypoint_kernel");
}

static int pomdp_launch(int64_t n, int32_t d, int32_t mode, float pomdp_prob, uint64_t seed, uint64_t step, const uint64_t* step_ptr,
                        int64_t env_id_base, int32_t stream_id, const float* in, float* out, void* stream);

extern "C" int ozl_pomdp_observation(int64_t n, int32_t d, int32_t mode, float pomdp_prob, uint64_t seed, uint64_t step,
                                     int64_t env_id_base, int32_t stream_id, const float* in, float* out, void* stream) {
    return pomdp_launch(n, d, mode, pomdp_prob, seed, step, nullptr, env_id_base, stream_id, in, out, stream);
}

extern "C" int ozl_pomdp_observation_dev(int64_t n, int32_t d, int32_t mode, float pomdp_prob, uint64_t seed,
                                         const uint64_t* step_ptr, int64_t env_id_base, int32_t stream_id, const float* in,
                                         float* out, void* stream) {
    if (!step_ptr) return set_error("ozl_pomdp_observation_dev: step_ptr is NULL");
    return pomdp_launch(n, d, mode, pomdp_prob, seed, 0, step_ptr, env_id_base, stream_id, in, out, stream);
}

static int pomdp_launch(int64_t n, int32_t d, int32_t mode, float pomdp_prob, uint64_t seed, uint64_t step, const uint64_t* step_ptr,
                        int64_t env_id_base, int32_t stream_id, const float* in, float* out, void* stream) {
    OZL_N_CHECK("ozl_pomdp_observation");
    if (mode < 1 || mode > 3)
        return set_error("pomdp was not in ['flicker', 'random_noise', 'flickering_and_random_noise']!");   // POMDP.py:19-20
    if (d <= 0 || d > 64 || !in || !out) return set_error("ozl_pomdp_observation: bad arguments");
    const float flick = (mode == 3) ? 0.1f : pomdp_prob;                                   // POMDP.py:16-18
    const float lo = (float)(1.0 - (double)pomdp_prob), hi = (float)(1.0 + (double)pomdp_prob);
    pomdp_kernel<<<nblk(n, 256), 256, 0, st>>>(n, d, mode, flick, lo, hi - lo, seed, step, (uint32_t)env_id_base,
                                               (uint32_t)stream_id, in, out, (const unsigned long long*)step_ptr);
    return check_cuda(cudaGetLastError(), "pomdp_kernel");
}

extern "C" int ozl_episode_stats(int64_t n, const float* rew, const int64_t* done, float* ep_ret, int32_t* ep_len,
                                 float* ret_out, int32_t* len_out, void* stream) {
    OZL_N_CHECK("ozl_episode_stats");
    if (!rew || !done || !ep_ret || !ep_len || !ret_out || !len_out) return set_error("ozl_episode_stats: NULL buffer");
    episode_stats_kernel<<<nblk(n, 256), 256, 0, st>>>(n, rew, done, ep_ret, ep_len, ret_out, len_out);
    return check_cuda(cudaGetLastError(), "episode_stats_kernel");
}


extern "C" int ozl_noise_lambda_apply(int64_t n, int32_t width, float* tensor, const ozl_noise_lambda* spec, float clip, uint64_t seed,
                                      uint64_t step, const uint64_t* step_ptr, int64_t step_offset, uint64_t corr_epoch,
                                      int64_t env_id_base, int32_t which, void* stream) {
    if (!tensor || !spec) return set_error("ozl_noise_lambda_apply: NULL argument");
    if (n <= 0 || width <= 0 || width > 64) return set_error("ozl_noise_lambda_apply: n = %lld, width = %d (1..64)", (long long)n, width);
    if (spec->distribution != OZL_DR_GAUSSIAN && spec->distribution != OZL_DR_UNIFORM)
        return set_error("ozl_noise_lambda_apply: distribution must be gaussian or uniform (vec_task.py:586,618)");
    if (spec->operation != OZL_DR_ADDITIVE && spec->operation != OZL_DR_SCALING)
        return set_error("ozl_noise_lambda_apply: operation must be additive or scaling");
    if (which != 0 && which != 1) return set_error("ozl_noise_lambda_apply: which must be 0 (observations) or 1 (actions)");
    NoiseLambda s{spec->distribution, spec->operation, spec->a, spec->b, spec->a_corr, spec->b_corr};
    noise_lambda_kernel<<<nblk(n, 128), 128, 0, (cudaStream_t)stream>>>(
        n, width, tensor, s, clip, seed, step, reinterpret_cast<const unsigned long long*>(step_ptr), (long long)step_offset, corr_epoch,
        (uint32_t)env_id_base, (uint32_t)which);
    return check_cuda(cudaGetLastError(), "noise_lambda_kernel");
}
