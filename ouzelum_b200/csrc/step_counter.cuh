// Device-resident step counter of an env handle (graph-capturable: no step launch takes a host-changing argument).
//
// The counter is a 16-byte record of two uint64 words
//     word 0 = base | (shift << 58)        word 1 = units            step = base + (units >> shift)
// Every step launch retires exactly 2^shift work units in total (one per 128-env tile; block 0 also retires the padding
// up to the power of two), each block with ONE fire-and-forget reduction (`red.add`: no return value, no fence) after its
// threads have read the counter.  While a launch is in flight `units` therefore stays below the next multiple of 2^shift,
// so a block that starts late still reads the step the launch began with; once the launch has finished, the next launch
// on the stream reads step + 1.  There is no "last block" election: no returning atomic and no __threadfence sits at
// the tail of a CTA (measured on B200 with the election: 5.16 us per 16384-env step, 17.7 us at 262144 envs; without:
// 4.2 / 14.9 us).  The record is read with ONE 16-byte relaxed load; the step kernels let one thread per block read it
// and broadcast through shared memory, because every request lands on the same L2 slice.
#pragma once
#include <stdint.h>

namespace ozl {

constexpr int kStepShiftBit = 58;
constexpr unsigned long long kStepBaseMask = (1ull << kStepShiftBit) - 1ull;

__device__ __forceinline__ uint64_t read_step(const unsigned long long* w) {
    unsigned long long b, u;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(b), "=l"(u) : "l"(w));
    return (b & kStepBaseMask) + (u >> (b >> kStepShiftBit));
}

__device__ __forceinline__ void retire_units(unsigned long long* w, unsigned long long units) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(w + 1), "l"(units) : "memory");
}

__device__ __forceinline__ unsigned long long step_word0_dev(unsigned long long base, unsigned shift) {
    return (base & kStepBaseMask) | ((unsigned long long)shift << kStepShiftBit);
}
inline unsigned long long step_word0(unsigned long long base, unsigned shift) {
    return (base & kStepBaseMask) | ((unsigned long long)shift << kStepShiftBit);
}
inline unsigned long long step_from_words(const unsigned long long w[2]) {
    return (w[0] & kStepBaseMask) + (w[1] >> (w[0] >> kStepShiftBit));
}

}  // namespace ozl
