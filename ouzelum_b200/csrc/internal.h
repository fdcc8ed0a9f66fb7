// Internal (non-ABI) declarations shared by the translation units of libouzelum_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ouzelum_b200.h"
#include "quad_env.cuh"

namespace ozl {

int set_error(const char* fmt, ...);   // stores a thread-local message, returns 1
int check_cuda(cudaError_t e, const char* what);

constexpr int kMetricSlots = 32;       // metric accumulators are striped over 32 slots (256 B apart) to spread L2 atomics
constexpr int kMetricStride = 32;      // doubles per slot (16 used)

// Private SoA state of one handle.  Planes are float4-packed so one env == one 16-byte lane per plane
// and a warp touches 512 contiguous bytes per plane.
//   dynamic (read+written every step):  d0 {px,py,pz,qx} d1 {qy,qz,qw,vx} d2 {vy,vz,wx,wy} d3 {wz,T0,T1,T2} d4 {T3,ep_ret}
//   static  (read every step, written only on reset/resample):
//           s0 {tx,ty,tz,fault_eff} s1 {1/mass,ixx,iyy,izz} s2 {arm,thrust_scale,fault_word,mass}
struct Planes {
    float4 *d0, *d1, *d2, *d3;
    float2* d4;
    float4 *s0, *s1, *s2;
    unsigned long long* ctrl;   // [0] step counter, [1] block ticket
    double* metrics;            // kMetricSlots x kMetricStride
};

}  // namespace ozl

struct ozl_env {
    ozl_cfg cfg;
    ozl::DevCfg dev;
    ozl::Planes pl;
    int device;
    int sm_count;
    void* arena;                // single cudaMalloc backing all planes
    size_t arena_bytes;
};
