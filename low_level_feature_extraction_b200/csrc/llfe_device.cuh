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

// ---- palette noise (shared by k_palette.cu and k_fused.cu) -----------------------------
__device__ __forceinline__ uint32_t fmix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}

// Device-generated noise with the distribution of  int8(trunc(N(0, 0.5)))  :
// P(+-1) = 0.0227501 each, P(+-2) = 3.167e-5 each (|n| >= 3: 1e-9, dropped), quantised
// to 2^-21.  Counter-based: a function of (seed, pixel index) only, so every pass
// over the image regenerates the same noise.  NOT NumPy's MT19937 stream.
__device__ __forceinline__ int noise21(uint32_t u21) {
    // thresholds on a 21-bit uniform: [0,T1) -> +1, [T1,2T1) -> -1, [2T1,2T1+T2) -> +2, [..,2T1+2T2) -> -2
    constexpr uint32_t T1 = 47710;  // round(0.02275013 * 2^21)
    constexpr uint32_t T2 = 66;     // round(3.1671e-5 * 2^21)
    if (u21 >= 2 * T1 + 2 * T2) return 0;
    if (u21 < T1) return 1;
    if (u21 < 2 * T1) return -1;
    if (u21 < 2 * T1 + T2) return 2;
    return -2;
}

__device__ __forceinline__ void device_noise(uint64_t seed, uint64_t pix, int& nr, int& ng, int& nb) {
    uint32_t lo = (uint32_t)pix, hi = (uint32_t)(pix >> 32);
    uint32_t a = fmix32(lo * 0x9E3779B1u ^ (uint32_t)seed ^ (hi * 0x7F4A7C15u));
    uint32_t b = fmix32(a ^ (uint32_t)(seed >> 32) ^ 0x68E31DA4u);
    uint64_t r = ((uint64_t)a << 32) | b;
    nr = noise21((uint32_t)(r & 0x1fffff));
    ng = noise21((uint32_t)((r >> 21) & 0x1fffff));
    nb = noise21((uint32_t)((r >> 42) & 0x1fffff));
}

__device__ __forceinline__ uint32_t noisy_key(uint32_t b, uint32_t g, uint32_t r, int nr, int ng, int nb) {
    int R = min(max((int)r + nr, 0), 255), G = min(max((int)g + ng, 0), 255), B = min(max((int)b + nb, 0), 255);
    return ((uint32_t)R << 16) | ((uint32_t)G << 8) | (uint32_t)B;
}

