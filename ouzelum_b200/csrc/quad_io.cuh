// Plane I/O of the private env state (tiled SoA, internal.h) and the block epilogue (episode statistics + step-counter
// retirement) shared by every kernel that runs env_step(): quad_step.cu and the fused EKFLeeLanded step (ekf_lee_fused.cu).
#pragma once
#include "internal.h"
#include "bulk_copy.cuh"

namespace ozl {

// ------------------------------------------------------------------------------------------------ plane I/O
struct Loaded {
    float4 d0, d1, d2, d3, s0, s1, s2;
    float2 d4;
};

__device__ __forceinline__ void unpack(const Loaded& L, Env& e) {
    e.p[0] = L.d0.x; e.p[1] = L.d0.y; e.p[2] = L.d0.z;
    e.q[0] = L.d0.w; e.q[1] = L.d1.x; e.q[2] = L.d1.y; e.q[3] = L.d1.z;
    e.v[0] = L.d1.w; e.v[1] = L.d2.x; e.v[2] = L.d2.y;
    e.w[0] = L.d2.z; e.w[1] = L.d2.w; e.w[2] = L.d3.x;
    e.T[0] = L.d3.y; e.T[1] = L.d3.z; e.T[2] = L.d3.w; e.T[3] = L.d4.x;
    e.ep_ret = L.d4.y;
    e.tgt[0] = L.s0.x; e.tgt[1] = L.s0.y; e.tgt[2] = L.s0.z; e.eff = L.s0.w;
    e.mass = L.s1.x; e.ixx = L.s1.y; e.iyy = L.s1.z; e.izz = L.s1.w;
    e.arm = L.s2.x; e.ks = L.s2.y; e.fault = __float_as_uint(L.s2.z); e.km = L.s2.w;
}
__device__ __forceinline__ void load_env(const Planes& pl, int64_t i, Loaded& L) {
    L.d0 = *plane4_ptr(pl, 0, i); L.d1 = *plane4_ptr(pl, 1, i); L.d2 = *plane4_ptr(pl, 2, i); L.d3 = *plane4_ptr(pl, 3, i);
    L.d4 = *plane2_ptr(pl, i);
    L.s0 = *plane4_ptr(pl, 4, i); L.s1 = *plane4_ptr(pl, 5, i); L.s2 = *plane4_ptr(pl, 6, i);
}
__device__ __forceinline__ void store_dynamic(const Planes& pl, int64_t i, const Env& e) {
    *plane4_ptr(pl, 0, i) = make_float4(e.p[0], e.p[1], e.p[2], e.q[0]);
    *plane4_ptr(pl, 1, i) = make_float4(e.q[1], e.q[2], e.q[3], e.v[0]);
    *plane4_ptr(pl, 2, i) = make_float4(e.v[1], e.v[2], e.w[0], e.w[1]);
    *plane4_ptr(pl, 3, i) = make_float4(e.w[2], e.T[0], e.T[1], e.T[2]);
    *plane2_ptr(pl, i) = make_float2(e.T[3], e.ep_ret);
}
__device__ __forceinline__ void store_static(const Planes& pl, int64_t i, const Env& e) {
    *plane4_ptr(pl, 4, i) = make_float4(e.tgt[0], e.tgt[1], e.tgt[2], e.eff);
    *plane4_ptr(pl, 5, i) = make_float4(e.mass, e.ixx, e.iyy, e.izz);
    *plane4_ptr(pl, 6, i) = make_float4(e.arm, e.ks, __uint_as_float(e.fault), e.km);
}

// Episode statistics (K6) + step-counter retirement at the end of a block.
// Metrics: one warp-level reduction per warp (redux / shuffles), then lane 0 adds the non-zero sums straight into the
// warp's metric slot with fire-and-forget reductions (RED.ADD.F64): no shared memory, no block barrier, warps retire
// independently.  The per-episode quantities (return, length, time-out / crash causes) are only reduced in warps where
// an episode actually ended this step (warp-uniform branch on a ballot).
// Metric slots: [0] sum reward [1] sum episode return [2] landed episodes [8] env-steps [9] episodes [10] sum episode
// length [11] time-outs [12] crash(dist) [13] crash(z) [14] fault-active steps [15] resets applied.
template <int BLOCK>
__device__ __forceinline__ void block_epilogue(const DevCfg& c, const Planes& pl, bool valid, const StepOut& o,
                                               int n_here, unsigned long long units) {
    if (c.collect_metrics) {
        const unsigned full = 0xffffffffu;
        const int lane = threadIdx.x & 31;
        const bool done = valid && o.reset;
        double* m = pl.metrics + ((blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5)) % kMetricSlots) * kMetricStride;
        const double srew = warp_sum(valid ? (double)o.rew : 0.0);
        // three 1-bit flags per lane in 8-bit fields (<= 32 per field)
        const uint32_t pk = __reduce_add_sync(full, valid ? ((uint32_t)o.fault_active | ((uint32_t)o.did_reset << 8) |
                                                             ((uint32_t)o.landed_episode << 16)) : 0u);
        const unsigned done_mask = __ballot_sync(full, done);
        if (lane == 0) {
            if (srew != 0.0) atomicAdd(m + 0, srew);
            if (pk & 0xFFu) atomicAdd(m + 14, (double)(pk & 0xFFu));
            if ((pk >> 8) & 0xFFu) atomicAdd(m + 15, (double)((pk >> 8) & 0xFFu));
            if (pk >> 16) atomicAdd(m + 2, (double)(pk >> 16));
        }
        if (done_mask) {
            const double sret = warp_sum(done ? (double)o.ep_ret_done : 0.0);
            const int slen = __reduce_add_sync(full, done ? (int)o.prog : 0);
            const uint32_t pe = __reduce_add_sync(full, done ? ((uint32_t)o.timeout | ((uint32_t)o.crash_dist << 8) |
                                                                ((uint32_t)o.crash_z << 16)) : 0u);
            if (lane == 0) {
                if (sret != 0.0) atomicAdd(m + 1, sret);
                atomicAdd(m + 9, (double)__popc(done_mask));
                atomicAdd(m + 10, (double)slen);
                if (pe & 0xFFu) atomicAdd(m + 11, (double)(pe & 0xFFu));
                if ((pe >> 8) & 0xFFu) atomicAdd(m + 12, (double)((pe >> 8) & 0xFFu));
                if (pe >> 16) atomicAdd(m + 13, (double)(pe >> 16));
            }
        }
        if (threadIdx.x == 0 && n_here > 0) atomicAdd(m + 8, (double)n_here);
    }
    // the caller has passed a block barrier after every thread consumed its copy of the step index (step_counter.cuh)
    if (threadIdx.x == 0) retire_units(pl.ctrl, units);
}

}  // namespace ozl
