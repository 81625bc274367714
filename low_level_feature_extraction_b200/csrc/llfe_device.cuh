// Device-side helpers shared by the kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// 128-bit streaming load: read-only path, do not allocate in L1 (inputs are read once).
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream(const uint2* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// ---- BGR2GRAY, Q15:  Y = (3735 B + 19235 G + 9798 R + 16384) >> 15 -----------
__device__ __forceinline__ uint8_t gray_px(uint32_t b, uint32_t g, uint32_t r) {
    return (uint8_t)((3735u * b + 19235u * g + 9798u * r + 16384u) >> 15);
}

// Same value via doubled coefficients so the result sits in byte 2:
//   (7470 B + 38470 G + 19596 R + 32768) >> 16   (sum < 2^24)
// computed with IDP.2A (two u16 coefficients x two u8 pixels bytes per instruction).
#define LLFE_CB 7470u
#define LLFE_CG 38470u
#define LLFE_CR 19596u
#define LLFE_PK(lo, hi) ((uint32_t)(lo) | ((uint32_t)(hi) << 16))

// 4 pixels stored in 3 consecutive little-endian words (b0 g0 r0 b1 | g1 r1 b2 g2 | r2 b3 g3 r3);
// returns the 4 un-shifted sums t0..t3 (gray = t >> 16).
__device__ __forceinline__ void gray4_sums(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t& t0, uint32_t& t1,
                                           uint32_t& t2, uint32_t& t3) {
    t0 = __dp2a_lo(LLFE_PK(LLFE_CB, LLFE_CG), w0, 32768u);
    t0 = __dp2a_hi(LLFE_PK(LLFE_CR, 0), w0, t0);
    t1 = __dp2a_hi(LLFE_PK(0, LLFE_CB), w0, 32768u);
    t1 = __dp2a_lo(LLFE_PK(LLFE_CG, LLFE_CR), w1, t1);
    t2 = __dp2a_hi(LLFE_PK(LLFE_CB, LLFE_CG), w1, 32768u);
    t2 = __dp2a_lo(LLFE_PK(LLFE_CR, 0), w2, t2);
    t3 = __dp2a_lo(LLFE_PK(0, LLFE_CB), w2, 32768u);
    t3 = __dp2a_hi(LLFE_PK(LLFE_CG, LLFE_CR), w2, t3);
}

// 4 gray bytes packed little-endian (pixel 0 in byte 0)
__device__ __forceinline__ uint32_t gray4_packed(uint32_t w0, uint32_t w1, uint32_t w2) {
    uint32_t t0, t1, t2, t3;
    gray4_sums(w0, w1, w2, t0, t1, t2, t3);
    uint32_t lo = __byte_perm(t0, t1, 0x0062);  // [t0.b2, t1.b2, x, x]
    uint32_t hi = __byte_perm(t2, t3, 0x0062);
    return __byte_perm(lo, hi, 0x5410);
}

// BORDER_REFLECT_101 index (valid for any offset; n >= 1)
__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    int period = 2 * (n - 1);
    i %= period;
    if (i < 0) i += period;
    return i >= n ? period - i : i;
}
__device__ __forceinline__ int clampi(int i, int lo, int hi) { return min(max(i, lo), hi); }

__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
