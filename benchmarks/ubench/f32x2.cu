// microbenchmark: issue/pipe throughput of scalar FMUL/FADD vs packed FMUL2/FADD2 on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    const float2 m = make_float2(s, s * 1.0001f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = __fmul_rn(a[i].x, m.x); a[i].y = __fmul_rn(a[i].y, m.y); }
            if (MODE == 1) { a[i] = __fmul2_rn(a[i], m); }
            if (MODE == 2) { a[i].x = __fadd_rn(a[i].x, m.x); a[i].y = __fadd_rn(a[i].y, m.y); }
            if (MODE == 3) { a[i] = __fadd2_rn(a[i], m); }
            if (MODE == 4) { a[i].x = __fmaf_rn(a[i].x, m.x, m.y); a[i].y = __fmaf_rn(a[i].y, m.y, m.x); }
            if (MODE == 5) { a[i] = __ffma2_rn(a[i], m, m); }
        }
    }
    float acc = 0;
    for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, float* d) {
    const int iters = 4096, blocks = 148 * 4, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, 16, 1.0000001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters, 1.0000001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // scalar-equivalent FP ops per SM per clock
    double ops = (double)blocks * threads * iters * 16.0;     // 16 scalar results per iteration per thread
    double clk = 1.965e9;
    printf("%-10s %8.3f ms  %.1f scalar results / SM / clk (at 1.965 GHz)\n", name, ms, ops / (ms * 1e-3) / 148.0 / clk);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
    run<0>("FMUL x2", d); run<1>("FMUL2", d); run<2>("FADD x2", d); run<3>("FADD2", d); run<4>("FFMA x2", d); run<5>("FFMA2", d);
    return 0;
}
