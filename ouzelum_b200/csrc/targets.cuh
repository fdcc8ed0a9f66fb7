// K5 device code: waypoint-following ground vehicle that carries the landing target (see targets.cu for the reference map).
// Shared by husky_step_kernel (targets.cu) and the fused EKFLeeLanded step (ekf_lee_fused.cu).
#pragma once
#include "internal.h"


namespace ozl {

constexpr int kNumWaypoints = 100;           // landing.py:108
constexpr float kWheelBase = 0.54f, kWheelRadius = 0.165f, kMaxWheel = 15.0f;   // controllers.py:18-20
constexpr float kPi = 3.14159265358979323846f;

struct HuskyArgs {
    int64_t n;
    float4* pose;            // [n] x, y, heading, scale*direction
    int2* idx;               // [n] trajectory id (0 lemniscate, 1 circle, 2 square), waypoint index
    const float2* tables;    // [204] lemniscate(100) | circle(100) | square(4)
    const int64_t* reset;    // [n] drone reset flags (may be null)
    float* wheels;           // [n,4] right,left,right,left (may be null)
    float* target3;          // [n,3] landing target riding on the vehicle
    uint64_t seed, step;
    const unsigned long long* step_ptr;
    uint32_t env_id_base;
    float dt, thresh, kp_lin, kp_ang, ang_thresh, x_offset, target_z, respawn_limit;
};

__device__ __forceinline__ float map_to_pi(float a) {                   // controllers.py:5-13
    if (a > kPi) a -= 2.0f * kPi;
    if (a <= -kPi) a += 2.0f * kPi;
    return a;
}
__device__ __forceinline__ float2 lookup(const HuskyArgs& a, int traj, int index, float s) {
    const int len = (traj == 2) ? 4 : kNumWaypoints;
    const int k = index < len - 1 ? index : len - 1;
    const float2 w = __ldg(a.tables + traj * kNumWaypoints + k);
    return make_float2(w.x * s, w.y * s);
}
__device__ __forceinline__ void redraw(const HuskyArgs& a, uint32_t genv, uint64_t step, int& traj, float& s) {
    const uint4 r = draw_cold(a.seed, genv, step, P_HUSKY);
    traj = (int)__umulhi(r.x, 3u);                                           // torch.randint(0, 3)      landing.py:210/240
    const float scale = __fadd_rn(0.8f, __fmul_rn(0.4f, u01(r.y)));                              // rand*(1.2-0.8)+0.8       landing.py:211/241
    s = scale * ((r.z & 1u) ? 1.0f : -1.0f);                                 // randint(0,2)*2-1         landing.py:212/242
}

// One control step of the vehicle of env i; writes pose / waypoint index / wheel speeds / the landing target riding on the
// vehicle, and returns that target in `tgt_out` (for kernels that go on to use it from registers).
// (`p`, `id`: the vehicle's pose / trajectory record, loaded by the caller -- early, so that the DRAM round trip is off the path)
__device__ __forceinline__ void husky_step_env(const HuskyArgs& a, int64_t i, uint64_t step, float tgt_out[3], float4 p, int2 id) {
    const uint32_t genv = a.env_id_base + (uint32_t)i;
    // re-spawn a strayed vehicle when its drone resets (landing.py:263-270)
    if (a.reset && a.reset[i] != 0 && (fabsf(p.x) > a.respawn_limit || fabsf(p.y) > a.respawn_limit)) {
        const uint4 r = draw_cold(a.seed, genv, step, P_HUSKY + 1);
        p.x = __fadd_rn(__fmul_rn(3.0f, u01(r.x)), -1.5f);
        p.y = __fadd_rn(__fmul_rn(3.0f, u01(r.y)), -1.5f);
        p.z = 0.0f;
    }
    // waypoint state machine (landing.py:326-358)
    float2 tgt = lookup(a, id.x, id.y, p.w);
    float dx = tgt.x - p.x, dy = tgt.y - p.y;
    if (sqrtf(dx * dx + dy * dy) < a.thresh) id.y += 1;                       // :339-343
    if (id.y == kNumWaypoints || (id.x == 2 && id.y > 3)) {                   // :235-238
        redraw(a, genv, step, id.x, p.w);
        id.y = 0;
    }
    tgt = lookup(a, id.x, id.y, p.w);
    // differential_drive(pos, target, heading, (3.0, 1000))  (controllers.py:15-43)
    dx = tgt.x - p.x; dy = tgt.y - p.y;
    // atan2f(+0, dx) = 0 or pi; taken out of the library call because its internal division sees the zero numerator (see
    // div_rn_normal) whenever a vehicle drives along a line of constant y
    float bearing = dx < 0.0f ? kPi : 0.0f;
    if (dy != 0.0f) bearing = atan2f(dy, dx);
    float dth = map_to_pi(bearing - map_to_pi(p.z));
    if (dth < a.ang_thresh && dth > -a.ang_thresh) dth = 0.0f;
    const float lin = sqrtf(dx * dx + dy * dy) * a.kp_lin;
    const float ang = dth * a.kp_ang;
    float left = div_rn_normal(2.0f * lin + ang * kWheelBase, 2.0f * kWheelRadius);
    float right = div_rn_normal(2.0f * lin - ang * kWheelBase, 2.0f * kWheelRadius);
    const float mx = fmaxf(fabsf(left), fabsf(right));
    if (mx > kMaxWheel) { const float sc = kMaxWheel / mx; left *= sc; right *= sc; }
    if (a.wheels) reinterpret_cast<float4*>(a.wheels)[i] = make_float4(right, left, right, left);
    // kinematic unicycle in place of the PhysX vehicle
    const float v = kWheelRadius * (right + left) * 0.5f;
    const float wz = div_rn_normal(kWheelRadius * (left - right), kWheelBase);
    float sh, ch;
    sincosf(p.z, &sh, &ch);
    p.x = p.x + ch * (v * a.dt);
    p.y = p.y + sh * (v * a.dt);
    float h = p.z + wz * a.dt;
    h = h - (2.0f * kPi) * floorf(div_rn_normal(h, 2.0f * kPi));                          // heading in [0, 2 pi) like get_euler_xyz
    p.z = h;
    a.pose[i] = p;
    a.idx[i] = id;
    a.target3[i * 3] = p.x + a.x_offset;                                      // landing.py:373-374
    a.target3[i * 3 + 1] = p.y;
    a.target3[i * 3 + 2] = a.target_z;
    tgt_out[0] = p.x + a.x_offset; tgt_out[1] = p.y; tgt_out[2] = a.target_z;
}

__device__ __forceinline__ void husky_step_env(const HuskyArgs& a, int64_t i, uint64_t step, float tgt_out[3]) {
    husky_step_env(a, i, step, tgt_out, a.pose[i], a.idx[i]);
}

}  // namespace ozl

// host: validates and converts the C-ABI argument block (returns non-zero and sets the error message on failure)
int ozl_fill_husky_args(const ozl_husky_args* in, ozl::HuskyArgs& a, const char* who);
