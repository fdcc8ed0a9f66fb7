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
#include "targets.cuh"

namespace ozl {

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
    float tgt[3];
    husky_step_env(a, i, a.step_ptr ? read_step(a.step_ptr) : a.step, tgt);
}

}  // namespace ozl

using namespace ozl;

int ozl_fill_husky_args(const ozl_husky_args* in, ozl::HuskyArgs& a, const char* who) {
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
    if (ozl_fill_husky_args(in, a, "ozl_husky_init")) return 1;
    husky_init_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    return check_cuda(cudaGetLastError(), "husky_init_kernel");
}

extern "C" int ozl_husky_step(const ozl_husky_args* in, void* stream) {
    HuskyArgs a;
    if (ozl_fill_husky_args(in, a, "ozl_husky_step")) return 1;
    if (!a.tables) return set_error("ozl_husky_step: tables204x2 is NULL");
    husky_step_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    return check_cuda(cudaGetLastError(), "husky_step_kernel");
}
