// Shadow mask from the blurred gray plane (ShadowAnalyzer, shadow_analyzer pyc L17-24):
//   cv2.adaptiveThreshold(blurred, 255, GAUSSIAN_C, BINARY_INV, 11, 2) + the masked sum / count of `blurred`
// with OpenCV's exact float32 operation order: row pass s = k0*x[-5]; s = fma(x[i-5], k[i], s),
// column pass v = k5*r[0]; v = fma(r[+i] + r[-i], k[5+i], v), mean = rint(v), mask = (px - mean <= -2).
//
// The fused front kernel (k_fused.cu) used to do this inline: 51 of its 162 thread-instructions per pixel,
// scalar FFMA chains squeezed between the edge and colour sections at 16 warps per SM.  Here the same
// arithmetic runs in a kernel of its own at ~24 instructions per pixel:
//   * a warp owns a 256-pixel span of a row band (lane L: 8 pixels, lanes 0/31 halo) and streams DOWN the rows,
//     TWO rows per step;
//   * row pass on VERTICAL pairs (row y, row y+1 of one pixel column) with packed fma.rn.f32x2: every tap of
//     the 11-tap chain finds its operand pair in one aligned register pair (a horizontal pairing would need a
//     second, shifted copy of the row for every other tap);
//   * column pass on HORIZONTAL pairs (pixels x, x+1 of one row) with add.rn.f32x2 / fma.rn.f32x2: no tap is
//     shifted in x, and the two output rows of a step share their ring loads (10 rows for 2 outputs);
//   * the 10-row window of row-pass results lives in a per-warp shared-memory ring whose slot numbers are
//     compile-time constants (the step is instantiated for the 5 phases of the ring);
//   * decision without integer conversion: bits(v + (1.5*2^23 - 2)) >= bits(1.5*2^23) + px  <=>  rint(v) >= px + 2;
//     masked sum with dp4a on the mask bytes, count with popc.
// f32x2 instructions are two IEEE round-to-nearest float32 operations: results are bit-identical to the scalar
// chain (tests/test_gpu_fused.py, test_gpu_ops.py::test_adaptive against cv2 itself).
//
// Requires W % 8 == 0, W >= 8 (other widths: k_adaptive in k_threshold.cu, which also knows OpenCV's tail columns).
#include "llfe_common.cuh"
#include "llfe_device.cuh"

namespace {

typedef unsigned long long u64;
constexpr unsigned FULL = 0xffffffffu;
constexpr int SWARPS = 4;
constexpr int S_OUT = 240;                 // 30 output lanes x 8 px per warp
constexpr int S_BAND_W = SWARPS * S_OUT;   // 960
constexpr int S_CTAS = 5;                  // resident CTAs per SM the kernel is compiled for (20 warps)
constexpr int SRING = 10;                  // rows vb-10 .. vb-1 of the row-pass results
constexpr int RING_F4 = SRING * 2 * 32;    // float4 per warp: [slot][half][lane]

#define SGK0 0x1.20c256p-7f
#define SGK1 0x1.bcb86ap-6f
#define SGK2 0x1.0ab50ap-4f
#define SGK3 0x1.f2464cp-4f
#define SGK4 0x1.6a7e1ep-3f
#define SGK5 0x1.9ac20ap-3f

__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 dup(float k) { return pk(k, k); }
__device__ __forceinline__ u64 shfl_up64(u64 v) { return __shfl_up_sync(FULL, v, 1); }
__device__ __forceinline__ u64 shfl_down64(u64 v) { return __shfl_down_sync(FULL, v, 1); }

// byte J (0..7) of the 8 pixels in (lo, hi) as the float bits of 2^23 + value
template <int J>
__device__ __forceinline__ float byte_as_biased_float(uint2 px) {
    return __uint_as_float(__byte_perm(J < 4 ? px.x : px.y, 0x4B000000u, 0x7440 + (J & 3)));
}
// byte J as the integer bits of 1.5 * 2^23 + value (the threshold word of the decision)
template <int J>
__device__ __forceinline__ uint32_t byte_as_threshold(uint2 px) {
    return __byte_perm(J < 4 ? px.x : px.y, 0x4B400000u, 0x7640 + (J & 3));   // [px, 0x00, 0x40, 0x4B]
}

struct ShadowArgs {
    const uint8_t* blurred;   // [n][h][w]
    uint8_t* mask;            // [n][h][w]
    unsigned long long* sum_count;  // [n][2] or null
    int h, w, rows_per_band;
};

// slot of window row d (d = -5 .. 4 relative to the first output row va) in ring phase PH (see k_shadow)
template <int PH, int D>
struct Slot {
    static constexpr int q = (D + 5) >> 1, sub = (D + 5) & 1;
    static constexpr int value = 2 * ((PH + q) % 5) + sub;
};

// horizontal pairs of one ring row half: (p0, p1) = pixels (0,1), (2,3) of the half
struct Row2 {
    u64 a, b;
};
template <int SLOT>
__device__ __forceinline__ Row2 ring_load(const float4* ring_lane, int half) {
    const float4 v = ring_lane[(SLOT * 2 + half) * 32];
    return Row2{pk(v.x, v.y), pk(v.z, v.w)};
}

template <int J>
__device__ __forceinline__ void decide(float v, uint2 centre, uint32_t& out_lo, uint32_t& out_hi) {
    // rint(v) >= px + 2  <=>  bits(v + (1.5 * 2^23 - 2)) >= bits(1.5 * 2^23) + px   (the magic constant is even, so the
    // addition rounds v to the nearest integer with ties to even exactly like rint; v is in [0, 255.001])
    const uint32_t fa = __float_as_uint(__fadd_rn(v, 12582910.0f));
    if (fa >= byte_as_threshold<J>(centre)) {
        if (J < 4) out_lo |= 0xffu << (8 * (J & 3));
        else out_hi |= 0xffu << (8 * (J & 3));
    }
}

// the two input rows (vb, vb + 1) of a step and the centre pixels of its output rows (vb - 5, vb - 4)
struct Rows {
    uint2 c0, c1, pc0, pc1;
};
__device__ __forceinline__ void load_rows(const uint8_t* __restrict__ img, int H, int W, int x, bool in_x, int vb, Rows& r) {
    r.c0 = r.c1 = r.pc0 = r.pc1 = make_uint2(0u, 0u);
    if (in_x) {
        r.c0 = *reinterpret_cast<const uint2*>(img + (size_t)clampi(vb, 0, H - 1) * W + x);
        r.c1 = *reinterpret_cast<const uint2*>(img + (size_t)clampi(vb + 1, 0, H - 1) * W + x);
        // the centre pixels were read five steps ago: cache hits (clamped: unused during the warm-up steps)
        r.pc0 = *reinterpret_cast<const uint2*>(img + (size_t)clampi(vb - 5, 0, H - 1) * W + x);
        r.pc1 = *reinterpret_cast<const uint2*>(img + (size_t)clampi(vb - 4, 0, H - 1) * W + x);
    }
}

// One step = rows (vb, vb + 1) in, rows (va, va + 1) = (vb - 5, vb - 4) out.  PH = (step index) % 5 fixes the ring slots.
template <int PH, bool PF>
__device__ __forceinline__ void shadow_step(const uint8_t* __restrict__ img, uint8_t* __restrict__ mout, int H, int W, int x,
                                            bool in_x, bool out_lane, bool left_edge, bool right_edge, int vb, bool emit0,
                                            bool emit1, float4* ring_lane, uint32_t& lsum, uint32_t& lcnt, Rows& next) {
    // ---- this step's rows were loaded one step ago; issue the next step's loads before any arithmetic
    //      (BORDER_REPLICATE in y by clamping, in x by copying the edge pixel into the halo lane)
    if (!PF) load_rows(img, H, W, x, in_x, vb, next);
    uint2 c0 = next.c0, c1 = next.c1;
    const uint2 pc0 = next.pc0, pc1 = next.pc1;
    if (PF) load_rows(img, H, W, x, in_x, vb + 2, next);
    if (left_edge) {    // lane 0 holds pixels -8..-1 := pixel 0
        const uint32_t a = __shfl_down_sync(FULL, c0.x, 1), b = __shfl_down_sync(FULL, c1.x, 1);
        if (x < 0) {
            c0.x = c0.y = (a & 0xffu) * 0x01010101u;
            c1.x = c1.y = (b & 0xffu) * 0x01010101u;
        }
    }
    if (right_edge) {   // the first lane beyond the image holds pixels W..W+7 := pixel W-1
        const uint32_t a = __shfl_up_sync(FULL, c0.y, 1), b = __shfl_up_sync(FULL, c1.y, 1);
        if (x == W) {
            c0.x = c0.y = (a >> 24) * 0x01010101u;
            c1.x = c1.y = (b >> 24) * 0x01010101u;
        }
    }
    // ---- exact floats as vertical pairs: P[g + 5] = (row vb pixel g, row vb+1 pixel g), g = -5 .. 12
    const u64 unbias = dup(-8388608.0f);
    u64 P[18];
#define CONV(J) P[5 + J] = add2(pk(byte_as_biased_float<J>(c0), byte_as_biased_float<J>(c1)), unbias);
    CONV(0) CONV(1) CONV(2) CONV(3) CONV(4) CONV(5) CONV(6) CONV(7)
#undef CONV
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        P[k] = shfl_up64(P[8 + k]);          // pixels -5..-1 = the previous lane's 3..7
        P[13 + k] = shfl_down64(P[5 + k]);   // pixels 8..12 = the next lane's 0..4
    }
    // ---- row pass, both rows at once
    const u64 K0 = dup(SGK0), K1 = dup(SGK1), K2 = dup(SGK2), K3 = dup(SGK3), K4 = dup(SGK4), K5 = dup(SGK5);
    float r0[8], r1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        u64 acc = mul2(K0, P[j]);
        acc = fma2(P[j + 1], K1, acc);
        acc = fma2(P[j + 2], K2, acc);
        acc = fma2(P[j + 3], K3, acc);
        acc = fma2(P[j + 4], K4, acc);
        acc = fma2(P[j + 5], K5, acc);
        acc = fma2(P[j + 6], K4, acc);
        acc = fma2(P[j + 7], K3, acc);
        acc = fma2(P[j + 8], K2, acc);
        acc = fma2(P[j + 9], K1, acc);
        acc = fma2(P[j + 10], K0, acc);
        upk(acc, r0[j], r1[j]);
    }
    // ---- column pass for rows va = vb - 5 (window rows va-5 .. va+5) and va + 1 (va-4 .. va+6):
    //      ring rows d = -5 .. 4, then r0 = row va + 5 and r1 = row va + 6 from registers
    if (emit0 && out_lane) {
        uint32_t o0_lo = 0, o0_hi = 0, o1_lo = 0, o1_hi = 0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const Row2 ca = ring_load<Slot<PH, 0>::value>(ring_lane, half), cb = ring_load<Slot<PH, 1>::value>(ring_lane, half);
            Row2 va2{mul2(K5, ca.a), mul2(K5, ca.b)}, vb2{mul2(K5, cb.a), mul2(K5, cb.b)};
            // tap 1: va: rows +1 (cb) and -1; vb: rows +2 and 0 (ca)
            const Row2 m1 = ring_load<Slot<PH, -1>::value>(ring_lane, half), p2 = ring_load<Slot<PH, 2>::value>(ring_lane, half);
            va2.a = fma2(add2(cb.a, m1.a), K4, va2.a);
            va2.b = fma2(add2(cb.b, m1.b), K4, va2.b);
            vb2.a = fma2(add2(p2.a, ca.a), K4, vb2.a);
            vb2.b = fma2(add2(p2.b, ca.b), K4, vb2.b);
            // tap 2: va: +2, -2; vb: +3, -1
            const Row2 m2 = ring_load<Slot<PH, -2>::value>(ring_lane, half), p3 = ring_load<Slot<PH, 3>::value>(ring_lane, half);
            va2.a = fma2(add2(p2.a, m2.a), K3, va2.a);
            va2.b = fma2(add2(p2.b, m2.b), K3, va2.b);
            vb2.a = fma2(add2(p3.a, m1.a), K3, vb2.a);
            vb2.b = fma2(add2(p3.b, m1.b), K3, vb2.b);
            // tap 3: va: +3, -3; vb: +4, -2
            const Row2 m3 = ring_load<Slot<PH, -3>::value>(ring_lane, half), p4 = ring_load<Slot<PH, 4>::value>(ring_lane, half);
            va2.a = fma2(add2(p3.a, m3.a), K2, va2.a);
            va2.b = fma2(add2(p3.b, m3.b), K2, va2.b);
            vb2.a = fma2(add2(p4.a, m2.a), K2, vb2.a);
            vb2.b = fma2(add2(p4.b, m2.b), K2, vb2.b);
            // tap 4: va: +4, -4; vb: +5 (r0), -3
            const Row2 m4 = ring_load<Slot<PH, -4>::value>(ring_lane, half);
            const Row2 p5{pk(r0[4 * half], r0[4 * half + 1]), pk(r0[4 * half + 2], r0[4 * half + 3])};
            va2.a = fma2(add2(p4.a, m4.a), K1, va2.a);
            va2.b = fma2(add2(p4.b, m4.b), K1, va2.b);
            vb2.a = fma2(add2(p5.a, m3.a), K1, vb2.a);
            vb2.b = fma2(add2(p5.b, m3.b), K1, vb2.b);
            // tap 5: va: +5 (r0), -5; vb: +6 (r1), -4
            const Row2 m5 = ring_load<Slot<PH, -5>::value>(ring_lane, half);
            const Row2 p6{pk(r1[4 * half], r1[4 * half + 1]), pk(r1[4 * half + 2], r1[4 * half + 3])};
            va2.a = fma2(add2(p5.a, m5.a), K0, va2.a);
            va2.b = fma2(add2(p5.b, m5.b), K0, va2.b);
            vb2.a = fma2(add2(p6.a, m4.a), K0, vb2.a);
            vb2.b = fma2(add2(p6.b, m4.b), K0, vb2.b);
            float a0, a1, a2, a3, b0, b1, b2, b3;
            upk(va2.a, a0, a1);
            upk(va2.b, a2, a3);
            upk(vb2.a, b0, b1);
            upk(vb2.b, b2, b3);
            if (half == 0) {
                decide<0>(a0, pc0, o0_lo, o0_hi);
                decide<1>(a1, pc0, o0_lo, o0_hi);
                decide<2>(a2, pc0, o0_lo, o0_hi);
                decide<3>(a3, pc0, o0_lo, o0_hi);
                decide<0>(b0, pc1, o1_lo, o1_hi);
                decide<1>(b1, pc1, o1_lo, o1_hi);
                decide<2>(b2, pc1, o1_lo, o1_hi);
                decide<3>(b3, pc1, o1_lo, o1_hi);
            } else {
                decide<4>(a0, pc0, o0_lo, o0_hi);
                decide<5>(a1, pc0, o0_lo, o0_hi);
                decide<6>(a2, pc0, o0_lo, o0_hi);
                decide<7>(a3, pc0, o0_lo, o0_hi);
                decide<4>(b0, pc1, o1_lo, o1_hi);
                decide<5>(b1, pc1, o1_lo, o1_hi);
                decide<6>(b2, pc1, o1_lo, o1_hi);
                decide<7>(b3, pc1, o1_lo, o1_hi);
            }
        }
        const int va = vb - 5;
        *reinterpret_cast<uint2*>(mout + (size_t)va * W + x) = make_uint2(o0_lo, o0_hi);
        lsum = __dp4a(pc0.x & o0_lo, 0x01010101u, lsum);
        lsum = __dp4a(pc0.y & o0_hi, 0x01010101u, lsum);
        lcnt += __popc(o0_lo) + __popc(o0_hi);    // 8 bits per selected pixel
        if (emit1) {
            *reinterpret_cast<uint2*>(mout + (size_t)(va + 1) * W + x) = make_uint2(o1_lo, o1_hi);
            lsum = __dp4a(pc1.x & o1_lo, 0x01010101u, lsum);
            lsum = __dp4a(pc1.y & o1_hi, 0x01010101u, lsum);
            lcnt += __popc(o1_lo) + __popc(o1_hi);
        }
    }
    // ---- rows vb, vb + 1 replace rows vb - 10, vb - 9 (slot pair PH)
    ring_lane[((2 * PH) * 2 + 0) * 32] = make_float4(r0[0], r0[1], r0[2], r0[3]);
    ring_lane[((2 * PH) * 2 + 1) * 32] = make_float4(r0[4], r0[5], r0[6], r0[7]);
    ring_lane[((2 * PH + 1) * 2 + 0) * 32] = make_float4(r1[0], r1[1], r1[2], r1[3]);
    ring_lane[((2 * PH + 1) * 2 + 1) * 32] = make_float4(r1[4], r1[5], r1[6], r1[7]);
}

template <bool PF, int CTAS>
__global__ void __launch_bounds__(SWARPS * 32, CTAS) k_shadow(ShadowArgs A) {
    extern __shared__ float4 s_ring[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int img = blockIdx.z, W = A.w, H = A.h;
    const int xs = blockIdx.x * S_BAND_W + warp * S_OUT - 8;   // first pixel of the warp span (lane 0 is halo)
    if (xs + 8 >= W) return;                                   // nothing to emit (warp-uniform)
    const int x = xs + 8 * lane;
    const bool in_x = x >= 0 && x < W;
    const bool out_lane = lane >= 1 && lane <= 30 && x < W;
    const bool left_edge = xs < 0, right_edge = xs + 256 > W;
    const int y0 = blockIdx.y * A.rows_per_band, y1 = min(H, y0 + A.rows_per_band);
    const uint8_t* src = A.blurred + (size_t)img * H * W;
    uint8_t* dst = A.mask + (size_t)img * H * W;
    float4* ring_lane = s_ring + (size_t)warp * RING_F4 + lane;
    uint32_t lsum = 0, lcnt = 0;
    // step t reads rows y0 - 5 + 2t (+1) and emits rows y0 + 2 (t - 5) (+1); 5 warm-up steps fill the ring
    const int steps = 5 + (y1 - y0 + 1) / 2;
    int vb = y0 - 5;
    Rows next;
    if (PF) load_rows(src, H, W, x, in_x, vb, next);
#define STEP(PH)                                                                                                       \
    {                                                                                                                  \
        const bool e0 = t >= 5 && vb - 5 < y1, e1 = e0 && vb - 4 < y1;                                                 \
        shadow_step<PH, PF>(src, dst, H, W, x, in_x, out_lane, left_edge, right_edge, vb, e0, e1, ring_lane, lsum, lcnt,   \
                        next);                                                                                         \
        vb += 2;                                                                                                       \
        if (++t >= steps) break;                                                                                       \
    }
    for (int t = 0;;) {
        STEP(0) STEP(1) STEP(2) STEP(3) STEP(4)
    }
#undef STEP
    if (A.sum_count) {
        lsum = warp_sum_u32(lsum);
        lcnt = warp_sum_u32(lcnt);
        if (lane == 0 && lcnt) {
            atomicAdd(&A.sum_count[2 * img], (unsigned long long)lsum);
            atomicAdd(&A.sum_count[2 * img + 1], (unsigned long long)(lcnt >> 3));
        }
    }
}

}  // namespace

bool shadow_split_supported(int h, int w) { return w % 8 == 0 && w >= 8 && h >= 1; }

// blurred (n, h, w) u8 -> mask (n, h, w) u8 in {0, 255} (+ sum / count of blurred under the mask), C = 2
int launch_shadow(llfe_ctx* ctx, const uint8_t* blurred, int n, int h, int w, uint8_t* mask, uint64_t* sum_count) {
    if (n == 0) return LLFE_OK;
    ShadowArgs A;
    A.blurred = blurred;
    A.mask = mask;
    A.sum_count = (unsigned long long*)sum_count;
    A.h = h;
    A.w = w;
    // Row bands: fill whole waves of the machine (S_CTAS CTAs per SM) with bands tall enough to amortise their
    // 10 warm-up rows.
    {
        const int xb = ceil_div(w, S_BAND_W);
        const int ctas_per_sm = ctx->shadow_variant == 2 ? 4 : ctx->shadow_variant == 3 ? 6 : S_CTAS;
        const double slots = (double)ctas_per_sm * (ctx->sm_count > 0 ? ctx->sm_count : 148);
        int best = 1;
        double best_score = -1.0;
        for (int bands = 1; bands <= 64 && (bands == 1 || h / bands >= 24); ++bands) {
            const int rpb = (ceil_div(h, bands) + 1) & ~1;   // even: a step emits two rows
            const double ctas = (double)xb * ceil_div(h, rpb) * n;
            const double waves = ctas / slots;
            const double fill = waves / (double)(long long)(waves + 0.999999);
            const double score = fill * rpb / (rpb + 10.0);
            if (score > best_score + 1e-9) {
                best_score = score;
                best = rpb;
            }
        }
        A.rows_per_band = best;
    }
    const size_t smem = (size_t)SWARPS * RING_F4 * sizeof(float4);   // 40 KB

    if (sum_count) LLFE_CUDA(cudaMemsetAsync(sum_count, 0, (size_t)n * 2 * sizeof(uint64_t), ctx->stream));
    dim3 grid(ceil_div(w, S_BAND_W), ceil_div(h, A.rows_per_band), n);
    LLFE_KERNEL(ctx, "k_shadow");
    switch (ctx->shadow_variant) {   // tuning variants (llfe_set_option "shadow_variant"); 0 is the default
        case 1: k_shadow<true, 5><<<grid, SWARPS * 32, smem, ctx->stream>>>(A); break;
        case 2: k_shadow<true, 4><<<grid, SWARPS * 32, smem, ctx->stream>>>(A); break;
        case 3: k_shadow<false, 6><<<grid, SWARPS * 32, smem, ctx->stream>>>(A); break;
        default: k_shadow<false, 5><<<grid, SWARPS * 32, smem, ctx->stream>>>(A); break;
    }
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}
