// Fused front end of the shapes + shadows pipeline: ONE read of the BGR image produces
//   EDGES  : the weak / strong Canny bit planes   (gray -> blur5 -> Sobel -> |dx|+|dy| -> NMS)
//   blurred: the blurred gray plane for the stand-alone shadow kernel (k_shadow.cu), optional
//   SHADOW : (round-1 layout, option "shadow_inline") the adaptive-threshold mask + masked sum/count inline
// with no gray / magnitude / NMS image in HBM.  The colour bitmap is a pointwise pass of its own
// (k_color_pass, k_palette.cu): inside this kernel its noise generator and bitmap lookups cost 1.25 ms per 256
// frames at 16 warps per SM, alone it is a streaming kernel.
//
// Mapping.  A warp owns a 256-pixel span of a row: lane L holds 8 consecutive pixels as four
// packed 16x2 registers, lanes 0 and 31 are halo (so a warp emits 240 pixels and a CTA of 8
// warps covers a 1920-pixel column band).  Warps stream DOWN the rows of their band keeping
// every vertical window in registers (5 h-blur rows, 3 blurred rows, 3 magnitude rows) and
// exchange horizontal neighbours with warp shuffles, so there is no block-level barrier in
// the main loop.  The 11-row float window of the adaptive threshold lives in a per-warp
// shared-memory ring (each lane reads back only what it wrote).
// Arithmetic is packed 16x2 integer (IDP.2A for the Q15 gray, VIMNMX.16x2 for |.|), exact.
//
// Requires W % 8 == 0 and W >= 8; other widths take the unfused kernels.
#include "llfe_common.cuh"
#include "llfe_device.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int WARPS = 4;
constexpr int OUT_PER_WARP = 240;            // 30 output lanes x 8 px
constexpr int BAND_W = WARPS * OUT_PER_WARP;  // 960
// resident CTAs per SM the kernel is compiled for: with the float ring (SHADOW) shared memory allows 4; the
// edges-only variant fits 96 registers at 5 CTAs (measured 8 % faster than 4 CTAs x 122 registers; 6 CTAs x 80
// registers spills and is 13 % slower)
constexpr int CTAS_FULL = 4, CTAS_EDGES = 5;
constexpr int RING = 10;   // rows vb-10 .. vb-1 of the 11-row window; the newest row is still in registers

// float32(cv2.getGaussianKernel(11, 0)), see k_threshold.cu
#define GK0 0x1.20c256p-7f
#define GK1 0x1.bcb86ap-6f
#define GK2 0x1.0ab50ap-4f
#define GK3 0x1.f2464cp-4f
#define GK4 0x1.6a7e1ep-3f
#define GK5 0x1.9ac20ap-3f

struct Q4 {
    uint32_t p0, p1, p2, p3;  // pixels (0,1) (2,3) (4,5) (6,7) as 16-bit lanes
};

__device__ __forceinline__ uint32_t mid(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x5432); }  // (a.hi, b.lo)
__device__ __forceinline__ uint32_t lo16(uint32_t a) { return a & 0xffffu; }
__device__ __forceinline__ uint32_t hi16(uint32_t a) { return a >> 16; }
__device__ __forceinline__ uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
__device__ __forceinline__ uint32_t vmin2(uint32_t a, uint32_t b) { return __vmins2(a, b); }

// pixel j (compile-time, -1..8) of a row held as Q4 + ext (lo = pixel -1, hi = pixel 8)
template <int J>
__device__ __forceinline__ uint32_t px(const Q4& q, uint32_t ext) {
    if (J == -1) return lo16(ext);
    if (J == 8) return hi16(ext);
    const uint32_t r = (J < 2) ? q.p0 : (J < 4) ? q.p1 : (J < 6) ? q.p2 : q.p3;
    return (J & 1) ? hi16(r) : lo16(r);
}

struct FusedArgs {
    const uint8_t* bgr;
    int n, h, w, rows_per_band;
    int low, high;
    uint32_t* weak;     // [n][h][wpr]
    uint32_t* strong;
    uint8_t* mask;      // [n][h][w] shadow mask
    unsigned long long* sum_count;  // [n][2]
    uint8_t* blur_out;  // [n][h][w] blurred gray plane for the stand-alone shadow kernel (k_shadow.cu), or null
};

// ---------------------------------------------------------------------------------------
// One lane's 8 pixels of a row as loaded: 24 bytes of BGR (+ 24 bytes of injected noise).
struct Raw {
    uint32_t w[6];
};

__device__ __forceinline__ void issue_row(const FusedArgs& A, int img, int y, int x, bool in_x, Raw& r) {
#pragma unroll
    for (int k = 0; k < 6; ++k) r.w[k] = 0u;
    if (in_x) {
        const size_t pix0 = (size_t)img * A.h * A.w + (size_t)y * A.w + x;
        const uint2* p = reinterpret_cast<const uint2*>(A.bgr + pix0 * 3);
        const uint2 a = p[0], b = p[1], c = p[2];
        r.w[0] = a.x; r.w[1] = a.y; r.w[2] = b.x; r.w[3] = b.y; r.w[4] = c.x; r.w[5] = c.y;
    }
}

__device__ __forceinline__ Q4 gray_of(const Raw& r) {
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7;
    gray4_sums(r.w[0], r.w[1], r.w[2], t0, t1, t2, t3);
    gray4_sums(r.w[3], r.w[4], r.w[5], t4, t5, t6, t7);
    Q4 g;
    g.p0 = __byte_perm(t0, t1, 0x7632);
    g.p1 = __byte_perm(t2, t3, 0x7632);
    g.p2 = __byte_perm(t4, t5, 0x7632);
    g.p3 = __byte_perm(t6, t7, 0x7632);
    return g;
}

// horizontal [1,4,6,4,1] on a gray row (16-bit lanes, max 4080)
__device__ __forceinline__ Q4 hblur(const Q4& g) {
    const uint32_t pm = __shfl_up_sync(FULL, g.p3, 1);    // pixels (-2,-1)
    const uint32_t pn = __shfl_down_sync(FULL, g.p0, 1);  // pixels (8,9)
    const uint32_t b0 = mid(pm, g.p0), b1 = mid(g.p0, g.p1), b2 = mid(g.p1, g.p2), b3 = mid(g.p2, g.p3),
                   b4 = mid(g.p3, pn);
    Q4 h;
    h.p0 = pm + g.p1 + 6u * g.p0 + 4u * (b0 + b1);
    h.p1 = g.p0 + g.p2 + 6u * g.p1 + 4u * (b1 + b2);
    h.p2 = g.p1 + g.p3 + 6u * g.p2 + 4u * (b2 + b3);
    h.p3 = g.p2 + pn + 6u * g.p3 + 4u * (b3 + b4);
    return h;
}

// horizontal [1,2,1] smoothing of a blurred row (needs the neighbours' edge pixels)
__device__ __forceinline__ Q4 hsmooth(const Q4& b) {
    const uint32_t pm = __shfl_up_sync(FULL, b.p3, 1), pn = __shfl_down_sync(FULL, b.p0, 1);
    const uint32_t l0 = mid(pm, b.p0), l1 = mid(b.p0, b.p1), l2 = mid(b.p1, b.p2), l3 = mid(b.p2, b.p3),
                   l4 = mid(b.p3, pn);
    Q4 t;
    t.p0 = l0 + 2u * b.p0 + l1;
    t.p1 = l1 + 2u * b.p1 + l2;
    t.p2 = l2 + 2u * b.p2 + l3;
    t.p3 = l3 + 2u * b.p3 + l4;
    return t;
}

struct Grad {
    Q4 ax, ay;       // |dx|, |dy| (<= 1020); bit 15 of each ax lane: (dx < 0) != (dy < 0)
};

// bit 15 of each 16-bit lane set iff a < b (values < 2^15)
__device__ __forceinline__ uint32_t lt2(uint32_t a, uint32_t b) { return (b | 0x80008000u) - a - 0x00010001u; }

__device__ __forceinline__ void sobel_row(const Q4& bm1, const Q4& b0, const Q4& bp1, const Q4& tm1, const Q4& tp1, Q4& mag,
                                          Grad& gr) {
    // vertical [1,2,1]
    Q4 s;
    s.p0 = bm1.p0 + 2u * b0.p0 + bp1.p0;
    s.p1 = bm1.p1 + 2u * b0.p1 + bp1.p1;
    s.p2 = bm1.p2 + 2u * b0.p2 + bp1.p2;
    s.p3 = bm1.p3 + 2u * b0.p3 + bp1.p3;
    const uint32_t pm = __shfl_up_sync(FULL, s.p3, 1), pn = __shfl_down_sync(FULL, s.p0, 1);
    const uint32_t l0 = mid(pm, s.p0), l1 = mid(s.p0, s.p1), l2 = mid(s.p1, s.p2), l3 = mid(s.p2, s.p3), l4 = mid(s.p3, pn);
    // dx = S(x+1) - S(x-1): right = l(k+1), left = l(k)
#define GRAD1(K, R, L, TP, TM)                                        \
    {                                                                 \
        const uint32_t ax = vmax2(R, L) - vmin2(R, L);                \
        const uint32_t ay = vmax2(TP, TM) - vmin2(TP, TM);            \
        gr.ax.K = ax | ((lt2(R, L) ^ lt2(TP, TM)) & 0x80008000u);     \
        gr.ay.K = ay;                                                 \
        mag.K = ax + ay;                                              \
    }
    GRAD1(p0, l1, l0, tp1.p0, tm1.p0)
    GRAD1(p1, l2, l1, tp1.p1, tm1.p1)
    GRAD1(p2, l3, l2, tp1.p2, tm1.p2)
    GRAD1(p3, l4, l3, tp1.p3, tm1.p3)
#undef GRAD1
}

__device__ __forceinline__ uint32_t ext_of(const Q4& m) {
    const uint32_t pm = __shfl_up_sync(FULL, m.p3, 1), pn = __shfl_down_sync(FULL, m.p0, 1);
    return hi16(pm) | (pn << 16);  // lo = pixel -1, hi = pixel 8
}

// non-maximum suppression of one pixel (scalar; only reached where mag > low)
template <int J>
__device__ __forceinline__ void nms_px(const Q4& m0, uint32_t e0, const Q4& m1, uint32_t e1, const Q4& m2, uint32_t e2,
                                       const Grad& gr, int low, int high, uint32_t& wbits, uint32_t& sbits) {
    const int m = (int)px<J>(m1, e1);
    if (m <= low) return;
    const uint32_t axs = px<J>(gr.ax, 0u);   // |dx| with the sign-difference flag in bit 15
    const int ax = (int)(axs & 0x7fffu), ay = (int)px<J>(gr.ay, 0u) << 15;
    const int tg22x = ax * 13573;
    bool keep;
    if (ay < tg22x) {
        keep = (m > (int)px<J - 1>(m1, e1)) && (m >= (int)px<J + 1>(m1, e1));
    } else if (ay > tg22x + (ax << 16)) {
        keep = (m > (int)px<J>(m0, e0)) && (m >= (int)px<J>(m2, e2));
    } else {
        const bool neg = (axs & 0x8000u) != 0;
        keep = neg ? ((m > (int)px<J + 1>(m0, e0)) && (m > (int)px<J - 1>(m2, e2)))
                   : ((m > (int)px<J - 1>(m0, e0)) && (m > (int)px<J + 1>(m2, e2)));
    }
    if (keep) {
        wbits |= 1u << J;
        if (m > high) sbits |= 1u << J;
    }
}

__device__ __forceinline__ float u2f(uint32_t v) { return __uint_as_float(0x4b000000u | v) - 8388608.0f; }

template <bool EDGES, bool SHADOW>
__global__ void __launch_bounds__(WARPS * 32, SHADOW ? CTAS_FULL : CTAS_EDGES) k_fused(FusedArgs A) {
    // dynamic shared memory, SHADOW only, per warp: [RING][2][32] float4 (row-pass results) + [RING][32] uint2 (the
    // blurred pixels as bytes) = 1280 B per row: 4 CTAs per SM
    extern __shared__ float4 dyn_smem[];
    float4* ring_all = dyn_smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int img = blockIdx.z;
    const int W = A.w, H = A.h;
    const int xs = blockIdx.x * BAND_W + warp * OUT_PER_WARP - 8;  // first pixel of the warp span
    if (xs + 8 >= W) return;                                       // nothing to emit (warp-uniform)
    const int x = xs + 8 * lane;
    const bool in_x = x >= 0 && x < W;
    const bool out_lane = lane >= 1 && lane <= 30 && x < W;
    const int y0 = blockIdx.y * A.rows_per_band, y1 = min(H, y0 + A.rows_per_band);
    const bool left_edge = xs < 0;              // warp-uniform: lane 0 is outside the image
    const bool right_edge = xs + 256 > W;       // warp-uniform: some lanes are outside the image
    const int wpr = (W + 31) >> 5;
    float4* ring = SHADOW ? ring_all + (size_t)warp * (RING * 2 * 32 + RING * 16) : nullptr;
    uint2* ring_px = SHADOW ? reinterpret_cast<uint2*>(ring + RING * 2 * 32) : nullptr;

    const int vb_lo = y0 - (SHADOW ? 5 : 2);
    const int vb_hi = y1 - 1 + (SHADOW ? 5 : 2);

    // Vertical [1,4,6,4,1] in transposed form: s0..s3 are the partial sums of the next four outputs (s0 lacks only the
    // newest row), so a new h-blurred row costs one multiply-add per accumulator and nothing has to be shifted along
    // a window of rows.  The rounding constant rides in through s3.
    Q4 s0 = {0u, 0u, 0u, 0u}, s1 = s0, s2 = s0, s3 = s0;
    Q4 bw[3], tw[3];           // blurred rows vb-2, vb-1, vb and their [1,2,1] smoothing
    Q4 mw[3];                  // magnitude rows vs-2, vs-1, vs
    uint32_t me[3] = {0u, 0u, 0u};
    Grad gprev, gcur;          // gradient of rows vs-1 (the NMS row) and vs
    bw[0] = bw[1] = bw[2] = tw[0] = tw[1] = tw[2] = mw[0] = mw[1] = mw[2] = Q4{0u, 0u, 0u, 0u};
    gprev.ax = gprev.ay = gcur.ax = gcur.ay = Q4{0u, 0u, 0u, 0u};
    Q4 blurred = {0u, 0u, 0u, 0u};
    uint32_t lsum = 0, lcnt = 0;
    int ring_slot = RING - 1;
    int rb_prev = 0;
    bool primed = false;

    // Row feed with one row of prefetch: the words of gray row `next_vy` are already in flight
    // while the previous row is being processed.
    const int vy_last = clampi(vb_hi, 0, H - 1) + 2;
    int next_vy = clampi(vb_lo, 0, H - 1) - 2;
    Raw raw_next;
    issue_row(A, img, reflect101_near(next_vy, H), x, in_x, raw_next);
    auto gray_row = [&]() -> Q4 {
        const Raw cur = raw_next;
        next_vy++;
        if (next_vy <= vy_last) issue_row(A, img, reflect101_near(next_vy, H), x, in_x, raw_next);
        // virtual gray row vy: BORDER_REFLECT_101 in y (done by the loader); in x the halo lanes are patched here
        Q4 g = gray_of(cur);
        if (left_edge) {   // lane 0 holds pixels -8..-1: (-2,-1) := (2,1)
            const uint32_t n0 = __shfl_down_sync(FULL, g.p0, 1), n1 = __shfl_down_sync(FULL, g.p1, 1);
            if (x == -8) g.p3 = lo16(n1) | (n0 & 0xffff0000u);
        }
        if (right_edge) {  // first lane beyond the image holds pixels W..W+7: (W, W+1) := (W-2, W-3)
            const uint32_t q2 = __shfl_up_sync(FULL, g.p2, 1), q3 = __shfl_up_sync(FULL, g.p3, 1);
            if (x == W) g.p0 = lo16(q3) | (q2 & 0xffff0000u);
        }
        return g;
    };

    for (int vb = vb_lo; vb <= vb_hi; ++vb) {
        const int rb = clampi(vb, 0, H - 1);
        if (!primed || rb != rb_prev) {
#define VB_FEED(P)                                   \
    {                                                \
        const uint32_t h_ = hrow.P;                  \
        const uint32_t v_ = s0.P + h_;               \
        s0.P = s1.P + 4u * h_;                       \
        s1.P = s2.P + 6u * h_;                       \
        s2.P = s3.P + 4u * h_;                       \
        s3.P = h_ + 0x00800080u;                     \
        blurred.P = __byte_perm(v_, 0u, 0x4341);     \
    }
            if (!primed) {
                primed = true;
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // rows rb-2 .. rb+1 only fill the accumulators
                    const Q4 hrow = hblur(gray_row());
                    VB_FEED(p0) VB_FEED(p1) VB_FEED(p2) VB_FEED(p3)
                }
            }
            {
                const Q4 hrow = hblur(gray_row());
                VB_FEED(p0) VB_FEED(p1) VB_FEED(p2) VB_FEED(p3)
            }
#undef VB_FEED
            rb_prev = rb;
            // BORDER_REPLICATE in x for the consumers of the blurred image
            if (left_edge) {
                const uint32_t v = lo16(__shfl_down_sync(FULL, blurred.p0, 1)) * 0x00010001u;
                if (x < 0) blurred = Q4{v, v, v, v};
            }
            if (right_edge) {
                const uint32_t v = hi16(__shfl_up_sync(FULL, blurred.p3, 1)) * 0x00010001u;
                if (x == W) blurred = Q4{v, v, v, v};
            }
        }
        if (A.blur_out && vb >= y0 && vb < y1 && out_lane)   // warp-uniform except for the halo lanes
            *reinterpret_cast<uint2*>(A.blur_out + ((size_t)img * H + vb) * W + x) =
                make_uint2(__byte_perm(blurred.p0, blurred.p1, 0x6420), __byte_perm(blurred.p2, blurred.p3, 0x6420));
        if (EDGES) {
            bw[0] = bw[1];
            bw[1] = bw[2];
            bw[2] = blurred;
            tw[0] = tw[1];
            tw[1] = tw[2];
            tw[2] = hsmooth(blurred);
            const int vs = vb - 1;  // Sobel row (needs blurred rows vs-1, vs, vs+1)
            if (vb >= vb_lo + 2) {
                mw[0] = mw[1];
                mw[1] = mw[2];
                me[0] = me[1];
                me[1] = me[2];
                gprev = gcur;
                Q4 mag;
                sobel_row(bw[0], bw[1], bw[2], tw[0], tw[2], mag, gcur);
                if (vs < 0 || vs >= H || !in_x) mag = Q4{0u, 0u, 0u, 0u};  // zero magnitude outside the image
                mw[2] = mag;
                me[2] = ext_of(mag);
                const int vn = vs - 1;  // NMS row
                if (vn >= y0 && vn < y1) {
                    uint32_t wbits = 0, sbits = 0;
                    const uint32_t thr = (0x7fffu - (uint32_t)A.low) * 0x00010001u;  // bit 15 <=> mag > low
                    const uint32_t any = ((mw[1].p0 + thr) | (mw[1].p1 + thr) | (mw[1].p2 + thr) | (mw[1].p3 + thr)) & 0x80008000u;
                    if (any) {
                        nms_px<0>(mw[0], me[0], mw[1], me[1], mw[2], me[2], gprev, A.low, A.high, wbits, sbits);
                        nms_px<1>(mw[0], me[0], mw[1], me[1], mw[2], me[2], gprev, A.low, A.high, wbits, sbits);
                        nms_px<2>(mw[0], me[0], mw[1], me[1], mw[2], me[2], gprev, A.low, A.high, wbits, sbits);
                        nms_px<3>(mw[0], me[0], mw[1], me[1], mw[2], me[2], gprev, A.low, A.high, wbits, sbits);
                        nms_px<4>(mw[0], me[0], mw[1], me[1], mw[2], me[2], gprev, A.low, A.high, wbits, sbits);
                        nms_px<5>(mw[0], me[0], mw[1], me[1], mw[2], me[2], gprev, A.low, A.high, wbits, sbits);
                        nms_px<6>(mw[0], me[0], mw[1], me[1], mw[2], me[2], gprev, A.low, A.high, wbits, sbits);
                        nms_px<7>(mw[0], me[0], mw[1], me[1], mw[2], me[2], gprev, A.low, A.high, wbits, sbits);
                    }
                    // one byte of each plane per lane (row padding beyond W is cleared by the launcher)
                    if (out_lane) {
                        const size_t o = ((size_t)img * H + vn) * wpr * 4 + (x >> 3);
                        reinterpret_cast<uint8_t*>(A.weak)[o] = (uint8_t)wbits;
                        reinterpret_cast<uint8_t*>(A.strong)[o] = (uint8_t)sbits;
                    }
                }
            }
        }
        if (SHADOW) {
            // row pass of the 11x11 Gaussian on the blurred row: s = k0*x[-5]; s = fma(x[i-5], k[i], s)
            float f[18];
            f[5] = u2f(lo16(blurred.p0));
            f[6] = u2f(hi16(blurred.p0));
            f[7] = u2f(lo16(blurred.p1));
            f[8] = u2f(hi16(blurred.p1));
            f[9] = u2f(lo16(blurred.p2));
            f[10] = u2f(hi16(blurred.p2));
            f[11] = u2f(lo16(blurred.p3));
            f[12] = u2f(hi16(blurred.p3));
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                f[k] = __shfl_up_sync(FULL, f[8 + k], 1);      // pixels -5..-1 = previous lane's 3..7
                f[13 + k] = __shfl_down_sync(FULL, f[5 + k], 1);  // pixels 8..12 = next lane's 0..4
            }
            float r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float acc = __fmul_rn(GK0, f[j]);
                acc = __fmaf_rn(f[j + 1], GK1, acc);
                acc = __fmaf_rn(f[j + 2], GK2, acc);
                acc = __fmaf_rn(f[j + 3], GK3, acc);
                acc = __fmaf_rn(f[j + 4], GK4, acc);
                acc = __fmaf_rn(f[j + 5], GK5, acc);
                acc = __fmaf_rn(f[j + 6], GK4, acc);
                acc = __fmaf_rn(f[j + 7], GK3, acc);
                acc = __fmaf_rn(f[j + 8], GK2, acc);
                acc = __fmaf_rn(f[j + 9], GK1, acc);
                acc = __fmaf_rn(f[j + 10], GK0, acc);
                r[j] = acc;
            }
            // ring slot of row vb: a running counter (no modulo in the loop).  The slot still holds row vb - 10
            // (the oldest row of this step's window); it is overwritten after the column pass.
            ring_slot = ring_slot + 1 == RING ? 0 : ring_slot + 1;
            const int slot = ring_slot;
            const int va = vb - 5;
            if (vb >= vb_lo + 10 && va >= y0 && va < y1 && out_lane) {
                float v[8];
                // row va + d (d = -5 .. 4) sits at slot + 5 + d (mod RING); row va + 5 = vb is r[] itself
                auto row_at = [&](int d, float* o) {
                    int s = slot + 5 + d;
                    if (s >= RING) s -= RING;
                    const float4* p = ring + (size_t)s * 2 * 32 + lane;
                    const float4 a = p[0], b = p[32];
                    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
                };
                float c0[8];
                row_at(0, c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __fmul_rn(GK5, c0[j]);
                const float kk[5] = {GK4, GK3, GK2, GK1, GK0};
#pragma unroll
                for (int i = 1; i <= 5; ++i) {
                    float up[8], dn[8];
                    if (i < 5) {
                        row_at(i, dn);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) dn[j] = r[j];
                    }
                    row_at(-i, up);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __fmaf_rn(__fadd_rn(dn[j], up[j]), kk[i - 1], v[j]);
                }
                int sc = slot + 5;
                if (sc >= RING) sc -= RING;
                const uint2 cb = ring_px[(size_t)sc * 32 + lane];   // 8 blurred pixels of row va, one byte each
                uint32_t out_lo = 0, out_hi = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // rint (half to even) via the 1.5*2^23 trick; v is in [0, 255.x]
                    const int mean = min(max((int)(__float_as_uint(__fadd_rn(v[j], 12582912.0f)) & 0x7fffffu) - 0x400000, 0), 255);
                    const int pxv = (int)(((j < 4 ? cb.x : cb.y) >> (8 * (j & 3))) & 0xffu);
                    const bool on = (pxv - mean) <= -2;
                    if (on) {
                        lsum += pxv;
                        lcnt += 1;
                        if (j < 4) out_lo |= 0xffu << (8 * j);
                        else out_hi |= 0xffu << (8 * (j - 4));
                    }
                }
                *reinterpret_cast<uint2*>(A.mask + ((size_t)img * H + va) * W + x) = make_uint2(out_lo, out_hi);
            }
            // now row vb replaces row vb - 10
            float4* rs = ring + (size_t)slot * 2 * 32 + lane;
            rs[0] = make_float4(r[0], r[1], r[2], r[3]);
            rs[32] = make_float4(r[4], r[5], r[6], r[7]);
            ring_px[(size_t)slot * 32 + lane] = make_uint2(__byte_perm(blurred.p0, blurred.p1, 0x6420),
                                                           __byte_perm(blurred.p2, blurred.p3, 0x6420));
        }
    }
    if (SHADOW) {
        lsum = warp_sum_u32(lsum);
        lcnt = warp_sum_u32(lcnt);
        if (lane == 0 && lcnt) {
            atomicAdd(&A.sum_count[2 * img], (unsigned long long)lsum);
            atomicAdd(&A.sum_count[2 * img + 1], (unsigned long long)lcnt);
        }
    }
}

template <bool E, bool S>
int launch_v(llfe_ctx* ctx, const FusedArgs& A, dim3 grid, size_t smem) {
    if (smem > 48 * 1024 && llfe_first_use(ctx, (const void*)k_fused<E, S>))   // smem is a constant of the variant
        LLFE_CUDA(cudaFuncSetAttribute(k_fused<E, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LLFE_KERNEL(ctx, "k_fused");
    k_fused<E, S><<<grid, WARPS * 32, smem, ctx->stream>>>(A);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

}  // namespace

bool fused_supported(int h, int w) { return w % 8 == 0 && w >= 8 && h >= 1; }

// Any of weak/strong (both or none), mask (+sum_count) and blur_out may be null to disable that output.
int launch_fused(llfe_ctx* ctx, const uint8_t* bgr, int n, int h, int w, int low, int high, uint32_t* weak,
                 uint32_t* strong, uint8_t* mask, uint64_t* sum_count, uint8_t* blur_out) {
    FusedArgs A;
    A.blur_out = blur_out;
    A.bgr = bgr;
    A.n = n;
    A.h = h;
    A.w = w;
    A.low = low;
    A.high = high;
    A.weak = weak;
    A.strong = strong;
    A.mask = mask;
    A.sum_count = (unsigned long long*)sum_count;
    const bool E = weak != nullptr, S = mask != nullptr;
    if (S && !sum_count) {   // sums go somewhere even if the caller does not want them
        if (!ctx->dummy_sums) LLFE_CUDA(cudaMalloc(&ctx->dummy_sums, 65536 * 2 * sizeof(unsigned long long)));
        A.sum_count = ctx->dummy_sums;
    }
    // Row bands: the grid should fill whole waves of the machine (CTAS_* CTAs per SM) and the bands should be tall
    // enough to amortise their ~10 warm-up rows.  Pick the band count with the best product of the two.
    {
        const int xb = ceil_div(w, BAND_W);
        const double slots = (double)(S ? CTAS_FULL : CTAS_EDGES) * (ctx->sm_count > 0 ? ctx->sm_count : 148);
        const double halo = S ? 10.0 : 4.0;
        int best = 1;
        double best_score = -1.0;
        for (int bands = 1; bands <= 32 && (bands == 1 || h / bands >= 32); ++bands) {
            const int rpb = ceil_div(h, bands);
            const double ctas = (double)xb * ceil_div(h, rpb) * n;
            const double waves = ctas / slots;
            const double fill = waves / (double)(long long)(waves + 0.999999);
            const double score = fill * rpb / (rpb + halo);
            if (score > best_score + 1e-9) {
                best_score = score;
                best = bands;
            }
        }
        A.rows_per_band = ceil_div(h, best);
    }
    dim3 grid(ceil_div(w, BAND_W), ceil_div(h, A.rows_per_band), n);
    const size_t smem = S ? (size_t)WARPS * (RING * 2 * 32 + RING * 16) * sizeof(float4) : 0;
    if (S && sum_count) LLFE_CUDA(cudaMemsetAsync(sum_count, 0, (size_t)n * 2 * sizeof(uint64_t), ctx->stream));
    if (E && (w % 32)) {  // the kernel writes whole bytes of in-image pixels only: clear the padding bits
        const size_t pb = (size_t)n * h * plane_wpr(w) * sizeof(uint32_t);
        LLFE_CUDA(cudaMemsetAsync(weak, 0, pb, ctx->stream));
        LLFE_CUDA(cudaMemsetAsync(strong, 0, pb, ctx->stream));
    }
    if (E && S) return launch_v<true, true>(ctx, A, grid, smem);
    if (E) return launch_v<true, false>(ctx, A, grid, smem);
    if (S) return launch_v<false, true>(ctx, A, grid, smem);
    return launch_v<false, false>(ctx, A, grid, smem);   // blurred plane only
}
