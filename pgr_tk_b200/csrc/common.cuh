// common.cuh — shared device/host helpers of libpgr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/pgr_b200.h"

namespace pgr {

// thread-local last error (pgr_b200_last_error)
void set_error(const char *fmt, ...);
const char *get_error();
// caller-owned result buffers (pinned pool for large ones); released by pgr_b200_free
void *result_alloc(size_t bytes);
void result_free(void *p);

#define PGR_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            ::pgr::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return PGR_E_CUDA;                                                                  \
        }                                                                                       \
    } while (0)

// shmmrutils.rs:271-280 (Thomas Wang 64-bit mix).  Written with multiplies by constants so that ptxas maps the
// shift-adds to IMAD.WIDE on the FMA pipe and leaves the xor-shifts on the ALU pipe.
__host__ __device__ __forceinline__ uint64_t u64hash(uint64_t key) {
    key = key * 0x1FFFFFull - 1ull;   // (!key) + (key << 21)
    key ^= key >> 24;
    key *= 265ull;                    // (key + (key << 3)) + (key << 8)
    key ^= key >> 14;
    key *= 21ull;                     // (key + (key << 2)) + (key << 4)
    key ^= key >> 28;
    key *= 0x80000001ull;             // key + (key << 31)
    return key;
}

constexpr uint64_t HASH_XOR = 0xAD12CF59ull;  // shmmrutils.rs:491

// Device variant on split 32-bit words (same arithmetic; 64-bit constant multiplies spelled as IMAD.WIDE + IMAD).
#ifdef __CUDACC__
template <int S>
__device__ __forceinline__ void xorshift_r(uint32_t &lo, uint32_t &hi) {   // key ^= key >> S, 0 < S < 32
    const uint32_t t_lo = __funnelshift_r(lo, hi, S);
    const uint32_t t_hi = hi >> S;
    lo ^= t_lo; hi ^= t_hi;
}
// (hi:lo) * C + A (mod 2^64) with a 32-bit constant C: one IMAD.WIDE for the low word product and one IMAD for the
// high word (the compiler's generic 64-bit multiply spends three)
template <uint32_t C>
__device__ __forceinline__ void mul64c(uint32_t &lo, uint32_t &hi, uint64_t addend = 0) {
    const uint64_t t = (uint64_t)lo * (uint64_t)C + addend;
    hi = hi * C + (uint32_t)(t >> 32);
    lo = (uint32_t)t;
}
__device__ __forceinline__ void u64hash_dev32(uint32_t &lo, uint32_t &hi) {
    mul64c<0x1FFFFFu>(lo, hi, ~0ull);      // key * (2^21 - 1) - 1
    xorshift_r<24>(lo, hi);
    mul64c<265u>(lo, hi);
    xorshift_r<14>(lo, hi);
    mul64c<21u>(lo, hi);
    xorshift_r<28>(lo, hi);
    mul64c<0x80000001u>(lo, hi);
}
// as u64hash_dev32 with the -1 of the first step supplied by the caller (a kernel parameter, see L0Params::m1)
__device__ __forceinline__ void u64hash_dev32m(uint32_t &lo, uint32_t &hi, uint64_t m1) {
    mul64c<0x1FFFFFu>(lo, hi, m1);
    xorshift_r<24>(lo, hi);
    mul64c<265u>(lo, hi);
    xorshift_r<14>(lo, hi);
    mul64c<21u>(lo, hi);
    xorshift_r<28>(lo, hi);
    mul64c<0x80000001u>(lo, hi);
}
// PGR_L0_HIMASK: which right shifts of the HIGH word in the key loop's hashes run on the FMA pipe as IMAD.HI with an
// opaque power-of-two factor (a kernel parameter, so that the compiler cannot turn it back into a shift) instead of SHF on the
// ALU pipe, which is the pipe that bounds the kernel: bit 0 = the >> 24, bit 1 = the >> 14, bit 2 = the >> 28 of u64hash,
// bit 3 = the >> (64 - K) that makes the high words of the two plane registers.  A/B in profiles/r2_l0_kernel_himask_ab.txt.
#ifndef PGR_L0_HIMASK
#define PGR_L0_HIMASK 0
#endif
struct HiShift { uint32_t c24, c14, c28, chs; };   // 2^(32-24), 2^(32-14), 2^(32-28), 2^(32-(64-K))
template <int S, bool HI>
__device__ __forceinline__ void xorshift_r_x(uint32_t &lo, uint32_t &hi, uint32_t c) {   // key ^= key >> S, 0 < S < 32
    const uint32_t t_lo = __funnelshift_r(lo, hi, S);
    const uint32_t t_hi = HI ? __umulhi(hi, c) : (hi >> S);
    lo ^= t_lo; hi ^= t_hi;
}
__device__ __forceinline__ void u64hash_dev32x(uint32_t &lo, uint32_t &hi, uint64_t m1, const HiShift &c) {
    mul64c<0x1FFFFFu>(lo, hi, m1);
    xorshift_r_x<24, (PGR_L0_HIMASK & 1) != 0>(lo, hi, c.c24);
    mul64c<265u>(lo, hi);
    xorshift_r_x<14, (PGR_L0_HIMASK & 2) != 0>(lo, hi, c.c14);
    mul64c<21u>(lo, hi);
    xorshift_r_x<28, (PGR_L0_HIMASK & 4) != 0>(lo, hi, c.c28);
    mul64c<0x80000001u>(lo, hi);
}
__device__ __forceinline__ uint64_t u64hash_dev(uint64_t key) {
    uint32_t lo = (uint32_t)key, hi = (uint32_t)(key >> 32);
    u64hash_dev32(lo, hi);
    return ((uint64_t)hi << 32) | lo;
}
#endif

// device-side mirror of pgr_shmmr_spec plus derived constants
struct SpecDev {
    uint32_t w, k, r, min_span, sketch;
};

// PGR_B200_TRACE=1: wall-clock stage trace on stderr (synchronises the device at every mark; debugging aid only)
void trace_mark(const char *what);

template <class T>
__host__ __device__ __forceinline__ T ceil_div(T a, T b) { return (a + b - 1) / b; }

}  // namespace pgr
