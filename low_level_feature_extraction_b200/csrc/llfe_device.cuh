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
// BORDER_REFLECT_101 for offsets that overshoot by less than the period (one mirror suffices); falls
// back to the general form for tiny n.  No integer division on the common path.
__device__ __forceinline__ int reflect101_near(int i, int n) {
    if (n < 16) return reflect101(i, n);
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

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

// fill `e` along runs of `w` inside one 32-bit word (both directions), e subset of w.
// Upward: adding the seeds to the word lets the carry ripple through each run from its lowest
// seed to the top of the run (and die in the gap above it); the bits the addition flipped, plus
// the seeds themselves, are the filled pixels.  Downward is the same on the bit-reversed words.
// 9 instructions instead of the 30 of a Kogge-Stone fill, and a much shorter dependency chain.
__device__ __forceinline__ uint32_t fill_up(uint32_t e, uint32_t w) { return (((w + e) ^ w) & w) | e; }
__device__ __forceinline__ uint32_t fill_word(uint32_t e, uint32_t w) {
    return fill_up(e, w) | __brev(fill_up(__brev(e), __brev(w)));
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

// Device-generated noise with the distribution of  int8(trunc(N(0, 0.5)))  per channel sample:
// P(non-zero) = p = 0.0455629 (+-1: 0.0227501 each, +-2: 3.167e-5 each; |n| >= 3 has 1e-9 and is
// dropped), independent across samples.  NOT NumPy's MT19937 stream.
//
// Non-zero samples are rare, so they are drawn SPARSELY by geometric skips: the number of zero
// samples before the next non-zero one is G = floor(log(u) / log(1 - p)) for uniform u, which
// reproduces independent Bernoulli(p) samples exactly (memorylessness) at one draw per non-zero
// sample instead of one per sample.  Every draw is a counter-based hash of (seed, image, block
// of 256 pixels, draw index).
#define LLFE_NOISE_INV_LOG2_Q (-14.863987f) /* 1 / log2(1 - 0.0455629) */

// draw #i of a block's stream: returns the skip (zeros before the next non-zero sample) and its value
__device__ __forceinline__ int noise_draw(uint32_t base, int i, int& value) {
    const uint32_t h = fmix32(base + (uint32_t)(i + 1) * 0x85EBCA77u);
    const float u = __fmul_rn(__fadd_rn((float)(h >> 9), 0.5f), 1.0f / 8388608.0f);  // (0, 1), 23 bits
    const int skip = (int)__fmul_rn(__log2f(u), LLFE_NOISE_INV_LOG2_Q);               // floor: the product is >= 0
    const uint32_t m = (h * 0x9E3779B1u) >> 16;                                   // 16 fresh-ish bits for the magnitude
    const int mag = 1 + (m < 91u);                                                // P(|n| = 2 | n != 0) = 1.39e-3
    value = (h & 0x100u) ? mag : -mag;                                            // bit 8: unused by u, sign
    return skip;
}

// Warp-cooperative generation for a BLOCK of 256 consecutive pixels (768 channel samples in memory order,
// linear pixel index >> 8): the 32 lanes draw 32 consecutive gaps of the block's geometric-skip stream at once,
// an inclusive warp scan turns the gaps into sample positions, and every lane applies its own hit to the block's
// bytes in shared memory.  The loop runs until the stream has passed the end of the block: 1.7 rounds on average,
// warp-uniform (a per-thread stream over 24 samples needs 2.1 draws on average but the warp waits for its slowest
// lane: ~5 rounds).  The noise is a pure function of (seed, image index, pixel position): any kernel that walks
// the blocks with this function regenerates the same values.
#define LLFE_NOISE_BLOCK_PX 256

__device__ __forceinline__ uint32_t noise_block_base(uint64_t seed, uint32_t image, uint32_t block) {
    uint32_t h = (uint32_t)seed * 0x9E3779B1u + ((uint32_t)(seed >> 32) ^ 0x7F4A7C15u) * 0x85EBCA6Bu;
    h = fmix32(h + image * 0x85EBCA77u);
    return fmix32(h ^ (block * 0xC2B2AE3Du));
}

// bytes: the block's samples in shared memory (byte k = channel sample k in memory order), n_valid <= 768 of them
// exist.  All 32 lanes call this together; __syncwarp() before (bytes written) and after (bytes read) is the caller's.
__device__ __forceinline__ void noise_apply_block(uint32_t base, uint8_t* bytes, int n_valid, int lane) {
    int carry = 0;   // samples the stream has passed so far
    for (int round = 0; carry < n_valid; ++round) {   // warp-uniform
        int v;
        int s = 1 + noise_draw(base, round * 32 + lane, v);   // gap to the next non-zero sample, inclusive
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        const int pos = carry + s - 1;
        if (pos < n_valid) {
            const int nv = (int)bytes[pos] + v;
            bytes[pos] = (uint8_t)min(max(nv, 0), 255);
        }
        carry += __shfl_sync(0xffffffffu, s, 31);
    }
}
