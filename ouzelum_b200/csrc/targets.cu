// K5: waypoint-following ground vehicle that carries the landing target, one env per thread.
// CPU twin: oracle/trajectories.py (HuskyFollower).  Reference:
//   isaacgymenvs/tasks/landing.py:208-213   per-env trajectory id / scale / direction / waypoint index
//   isaacgymenvs/tasks/landing.py:224-244   reset_completed_trajectories
//   isaacgymenvs/tasks/landing.py:319-364   set_husky_actions  (two O(N) Python loops with per-element branching)
//   isaacgymenvs/tasks/landing.py:263-270   re-spawn of a vehicle that strayed beyond 2 x envSpacing, on drone reset
//   isaacgymenvs/tasks/landing.py:373-374   target = vehicle.xy, x += 0.08 (ekf_lee_landed.py:628-629: -0.08), z = 0.377
//   isaacgymenvs/utils/controllers.py:5-43  map_to_pi / differential_drive
// The vehicle itself is a PhysX articulation in the reference; here the wheel speeds drive a kinematic unicycle
// (new, parity unpinned -- SURVEY 8a row G3).  Random re-draws come from the counter RNG (purpose P_HUSKY).
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
    const uint4 r = draw(a.seed, genv, step, P_HUSKY);
    traj = (int)__umulhi(r.x, 3u);                                           // torch.randint(0, 3)      landing.py:210/240
    const float scale = __fadd_rn(0.8f, __fmul_rn(0.4f, u01(r.y)));                              // rand*(1.2-0.8)+0.8       landing.py:211/241
    s = scale * ((r.z & 1u) ? 1.0f : -1.0f);                                 // randint(0,2)*2-1         landing.py:212/242
}

__global__ void husky_init_kernel(const HuskyArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    int traj; float s;
    redraw(a, a.env_id_base + (uint32_t)i, 1ull << 63, traj, s);
    a.pose[i] = make_float4(0.0f, 0.0f, 0.0f, s);
    a.idx[i] = make_int2(traj, 0);
    if (a.target3) { a.target3[i * 3] = a.x_offset; a.target3[i * 3 + 1] = 0.0f; a.target3[i * 3 + 2] = a.target_z; }
}

__global__ void __launch_bounds__(256)
husky_step_kernel(const HuskyArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint32_t genv = a.env_id_base + (uint32_t)i;
    const uint64_t step = a.step_ptr ? read_step(a.step_ptr) : a.step;
    float4 p = a.pose[i];
    int2 id = a.idx[i];
    // re-spawn a strayed vehicle when its drone resets (landing.py:263-270)
    if (a.reset && a.reset[i] != 0 && (fabsf(p.x) > a.respawn_limit || fabsf(p.y) > a.respawn_limit)) {
        const uint4 r = draw(a.seed, genv, step, P_HUSKY + 1);
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
    float dth = map_to_pi(atan2f(dy, dx) - map_to_pi(p.z));
    if (dth < a.ang_thresh && dth > -a.ang_thresh) dth = 0.0f;
    const float lin = sqrtf(dx * dx + dy * dy) * a.kp_lin;
    const float ang = dth * a.kp_ang;
    float left = (2.0f * lin + ang * kWheelBase) / (2.0f * kWheelRadius);
    float right = (2.0f * lin - ang * kWheelBase) / (2.0f * kWheelRadius);
    const float mx = fmaxf(fabsf(left), fabsf(right));
    if (mx > kMaxWheel) { const float sc = kMaxWheel / mx; left *= sc; right *= sc; }
    if (a.wheels) reinterpret_cast<float4*>(a.wheels)[i] = make_float4(right, left, right, left);
    // kinematic unicycle in place of the PhysX vehicle
    const float v = kWheelRadius * (right + left) * 0.5f;
    const float wz = kWheelRadius * (left - right) / kWheelBase;
    float sh, ch;
    sincosf(p.z, &sh, &ch);
    p.x = p.x + ch * (v * a.dt);
    p.y = p.y + sh * (v * a.dt);
    float h = p.z + wz * a.dt;
    h = h - (2.0f * kPi) * floorf(h / (2.0f * kPi));                          // heading in [0, 2 pi) like get_euler_xyz
    p.z = h;
    a.pose[i] = p;
    a.idx[i] = id;
    a.target3[i * 3] = p.x + a.x_offset;                                      // landing.py:373-374
    a.target3[i * 3 + 1] = p.y;
    a.target3[i * 3 + 2] = a.target_z;
}

}  // namespace ozl

using namespace ozl;

static int fill(const ozl_husky_args* in, HuskyArgs& a, const char* who) {
    if (!in) return set_error("%s: args is NULL", who);
    if (in->n <= 0) return set_error("%s: n must be > 0", who);
    if (!in->pose4 || !in->idx2 || !in->target3) return set_error("%s: NULL buffer", who);
    if (((uintptr_t)in->pose4 & 15) || ((uintptr_t)in->wheels4 & 15)) return set_error("%s: pose4/wheels4 must be 16-byte aligned", who);
    a.n = in->n; a.pose = (float4*)in->pose4; a.idx = (int2*)in->idx2; a.tables = (const float2*)in->tables204x2;
    a.reset = in->reset; a.wheels = in->wheels4; a.target3 = in->target3;
    a.seed = in->seed; a.step = in->step; a.step_ptr = (const unsigned long long*)in->step_ptr; a.env_id_base = (uint32_t)in->env_id_base;
    a.dt = in->dt; a.thresh = in->dist_thresh; a.kp_lin = in->kp_lin; a.kp_ang = in->kp_ang; a.ang_thresh = in->ang_thresh;
    a.x_offset = in->x_offset; a.target_z = in->target_z; a.respawn_limit = in->respawn_limit;
    return 0;
}

extern "C" int ozl_husky_init(const ozl_husky_args* in, void* stream) {
    HuskyArgs a;
    if (fill(in, a, "ozl_husky_init")) return 1;
    husky_init_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    return check_cuda(cudaGetLastError(), "husky_init_kernel");
}

extern "C" int ozl_husky_step(const ozl_husky_args* in, void* stream) {
    HuskyArgs a;
    if (fill(in, a, "ozl_husky_step")) return 1;
    if (!a.tables) return set_error("ozl_husky_step: tables204x2 is NULL");
    husky_step_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    return check_cuda(cudaGetLastError(), "husky_step_kernel");
}
