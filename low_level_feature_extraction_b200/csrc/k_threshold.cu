// Threshold masks:
//   adaptive : cv2.adaptiveThreshold(src,255,GAUSSIAN_C,BINARY_INV,11,C) with OpenCV's
//              exact float32 operation order (row pass: left-to-right FMA chain; column
//              pass: symmetric-pair FMA chain; BORDER_REPLICATE), optional masked sum/count.
//   otsu     : 256-bin histogram -> float64 sweep (first maximum wins) -> binarise,
//              optional "invert if mean(mask) > 127".
#include "llfe_common.cuh"
#include "llfe_device.cuh"

namespace {

// float32(cv2.getGaussianKernel(11, 0)): sigma = 2.0, symmetric (k[i] == k[10-i]).
// bit patterns 0x3c10612b 0x3cde5c35 0x3d855a85 0x3df92326 0x3e353f0f 0x3e4d6105
#define GK0 0x1.20c256p-7f
#define GK1 0x1.bcb86ap-6f
#define GK2 0x1.0ab50ap-4f
#define GK3 0x1.f2464cp-4f
#define GK4 0x1.6a7e1ep-3f
#define GK5 0x1.9ac20ap-3f

constexpr int ATW = 64;
constexpr int ATH = 32;

__global__ void __launch_bounds__(256) k_adaptive(const uint8_t* __restrict__ src, int h, int w, int C,
                                                  uint8_t* __restrict__ mask, unsigned long long* sum_count) {
    __shared__ uint8_t tin[ATH + 10][ATW + 10 + 2];
    __shared__ float rs[ATH + 10][ATW];
    __shared__ unsigned long long s_acc[2];
    const int img = blockIdx.z;
    const uint8_t* s = src + (size_t)img * h * w;
    uint8_t* m = mask + (size_t)img * h * w;
    const int x0 = blockIdx.x * ATW, y0 = blockIdx.y * ATH;
    const int tid = threadIdx.x;
    if (tid < 2) s_acc[tid] = 0;
    for (int i = tid; i < (ATH + 10) * (ATW + 10); i += 256) {
        int ry = i / (ATW + 10), rx = i - ry * (ATW + 10);
        int y = clampi(y0 - 5 + ry, 0, h - 1), x = clampi(x0 - 5 + rx, 0, w - 1);
        tin[ry][rx] = s[(size_t)y * w + x];
    }
    __syncthreads();
    // row pass: s = k0*x[-5]; s = fma(x[i-5], k[i], s) for i = 1..10 (left to right).
    // Tail columns (pinned against the installed cv2 binary, oracle/cvops.py:gauss11_f32): OpenCV's row
    // filter runs 8-wide and then 4-wide FMA vectors; the last w % 4 columns go through its scalar loop, whose
    // compiled code multiplies and adds separately for taps 1..8 and fuses only the last two taps.
    const int n8 = w & ~7, n4 = (w - n8 >= 4) ? n8 + 4 : n8;
    for (int i = tid; i < (ATH + 10) * ATW; i += 256) {
        int ry = i / ATW, rx = i - ry * ATW;
        const uint8_t* p = &tin[ry][rx];
        float acc = __fmul_rn(GK0, (float)p[0]);
        if (x0 + rx < n4) {
            acc = __fmaf_rn((float)p[1], GK1, acc);
            acc = __fmaf_rn((float)p[2], GK2, acc);
            acc = __fmaf_rn((float)p[3], GK3, acc);
            acc = __fmaf_rn((float)p[4], GK4, acc);
            acc = __fmaf_rn((float)p[5], GK5, acc);
            acc = __fmaf_rn((float)p[6], GK4, acc);
            acc = __fmaf_rn((float)p[7], GK3, acc);
            acc = __fmaf_rn((float)p[8], GK2, acc);
        } else {
            acc = __fadd_rn(__fmul_rn((float)p[1], GK1), acc);
            acc = __fadd_rn(__fmul_rn((float)p[2], GK2), acc);
            acc = __fadd_rn(__fmul_rn((float)p[3], GK3), acc);
            acc = __fadd_rn(__fmul_rn((float)p[4], GK4), acc);
            acc = __fadd_rn(__fmul_rn((float)p[5], GK5), acc);
            acc = __fadd_rn(__fmul_rn((float)p[6], GK4), acc);
            acc = __fadd_rn(__fmul_rn((float)p[7], GK3), acc);
            acc = __fadd_rn(__fmul_rn((float)p[8], GK2), acc);
        }
        acc = __fmaf_rn((float)p[9], GK1, acc);
        acc = __fmaf_rn((float)p[10], GK0, acc);
        rs[ry][rx] = acc;
    }
    __syncthreads();
    // column pass: v = k5*r[0]; v = fma(r[+i] + r[-i], k[5+i], v) for i = 1..5
    unsigned long long lsum = 0, lcnt = 0;
    for (int i = tid; i < ATH * ATW; i += 256) {
        int ry = i / ATW, rx = i - ry * ATW;
        int y = y0 + ry, x = x0 + rx;
        if (y >= h || x >= w) continue;
        float v = __fmul_rn(GK5, rs[ry + 5][rx]);
        if (x < n8) {
            v = __fmaf_rn(__fadd_rn(rs[ry + 6][rx], rs[ry + 4][rx]), GK4, v);
            v = __fmaf_rn(__fadd_rn(rs[ry + 7][rx], rs[ry + 3][rx]), GK3, v);
            v = __fmaf_rn(__fadd_rn(rs[ry + 8][rx], rs[ry + 2][rx]), GK2, v);
            v = __fmaf_rn(__fadd_rn(rs[ry + 9][rx], rs[ry + 1][rx]), GK1, v);
            v = __fmaf_rn(__fadd_rn(rs[ry + 10][rx], rs[ry + 0][rx]), GK0, v);
        } else {   // OpenCV's column filter: 8-wide FMA vectors, then separate multiply + add for the last w % 8 columns
            v = __fadd_rn(__fmul_rn(__fadd_rn(rs[ry + 6][rx], rs[ry + 4][rx]), GK4), v);
            v = __fadd_rn(__fmul_rn(__fadd_rn(rs[ry + 7][rx], rs[ry + 3][rx]), GK3), v);
            v = __fadd_rn(__fmul_rn(__fadd_rn(rs[ry + 8][rx], rs[ry + 2][rx]), GK2), v);
            v = __fadd_rn(__fmul_rn(__fadd_rn(rs[ry + 9][rx], rs[ry + 1][rx]), GK1), v);
            v = __fadd_rn(__fmul_rn(__fadd_rn(rs[ry + 10][rx], rs[ry + 0][rx]), GK0), v);
        }
        int mean = min(max(__float2int_rn(v), 0), 255);
        int px = tin[ry + 5][rx + 5];
        bool on = (px - mean) <= -C;
        m[(size_t)y * w + x] = on ? 255 : 0;
        if (on) {
            lsum += px;
            lcnt += 1;
        }
    }
    if (sum_count) {
        lsum = warp_sum_u64(lsum);
        lcnt = warp_sum_u64(lcnt);
        if ((tid & 31) == 0 && lcnt) {
            atomicAdd(&s_acc[0], lsum);
            atomicAdd(&s_acc[1], lcnt);
        }
        __syncthreads();
        if (tid == 0 && s_acc[1]) {
            atomicAdd(&sum_count[2 * img], s_acc[0]);
            atomicAdd(&sum_count[2 * img + 1], s_acc[1]);
        }
    }
}

// ------------------------------------------------------------------ otsu ---
__global__ void __launch_bounds__(256) k_hist256(const uint8_t* __restrict__ gray, size_t npix,
                                                 uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[8][256];
    const int img = blockIdx.y;
    const uint8_t* s = gray + (size_t)img * npix;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8 * 256; i += 256) (&sh[0][0])[i] = 0;
    __syncthreads();
    size_t stride = (size_t)gridDim.x * 256;
    size_t gid = blockIdx.x * (size_t)256 + tid;
    const bool vec = ((uintptr_t)s % 16) == 0;
    size_t nvec = vec ? npix / 16 : 0;
    for (size_t i = gid; i < nvec; i += stride) {
        uint4 v = ld_stream(reinterpret_cast<const uint4*>(s) + i);
        uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t x = wds[j];
            atomicAdd(&sh[warp][x & 255], 1u);
            atomicAdd(&sh[warp][(x >> 8) & 255], 1u);
            atomicAdd(&sh[warp][(x >> 16) & 255], 1u);
            atomicAdd(&sh[warp][x >> 24], 1u);
        }
    }
    for (size_t i = nvec * 16 + gid; i < npix; i += stride) atomicAdd(&sh[warp][s[i]], 1u);
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][tid];
    if (t) atomicAdd(&hist[(size_t)img * 256 + tid], t);
}

// One thread per image: the float64 Otsu sweep of OpenCV (getThreshVal_Otsu_8u).
__global__ void k_otsu_sweep(const uint32_t* __restrict__ hist, int n, size_t npix, int invert_if_light,
                             int32_t* __restrict__ thresh, int32_t* __restrict__ invert) {
    int img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= n) return;
    const uint32_t* hgram = hist + (size_t)img * 256;
    const double scale = 1.0 / (double)npix;
    double mu = 0.0;
    for (int i = 0; i < 256; ++i) mu = __dadd_rn(mu, __dmul_rn((double)i, (double)hgram[i]));
    mu = __dmul_rn(mu, scale);
    double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
    int max_val = 0;
    const double eps = 1.1920928955078125e-07;  // FLT_EPSILON, as in OpenCV
    for (int i = 0; i < 256; ++i) {
        double p_i = __dmul_rn((double)hgram[i], scale);
        mu1 = __dmul_rn(mu1, q1);
        q1 = __dadd_rn(q1, p_i);
        double q2 = __dsub_rn(1.0, q1);
        if (fmin(q1, q2) < eps || fmax(q1, q2) > 1.0 - eps) continue;
        mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
        double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
        double d = __dsub_rn(mu1, mu2);
        double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
        if (sigma > max_sigma) {
            max_sigma = sigma;
            max_val = i;
        }
    }
    thresh[img] = max_val;
    int inv = 0;
    if (invert_if_light) {
        unsigned long long on = 0;
        for (int i = max_val + 1; i < 256; ++i) on += hgram[i];
        // mean(mask) > 127  <=>  255 * on > 127 * npix  (exact)
        inv = (255ull * on > 127ull * (unsigned long long)npix) ? 1 : 0;
    }
    invert[img] = inv;
}

__global__ void __launch_bounds__(256) k_binarize(const uint8_t* __restrict__ gray, size_t npix,
                                                  const int32_t* __restrict__ thresh, const int32_t* __restrict__ invert,
                                                  uint8_t* __restrict__ mask) {
    const int img = blockIdx.y;
    const uint8_t* s = gray + (size_t)img * npix;
    uint8_t* m = mask + (size_t)img * npix;
    const uint32_t t = (uint32_t)thresh[img];
    const uint32_t flip = invert[img] ? 0xffffffffu : 0u;
    size_t stride = (size_t)gridDim.x * 256;
    size_t gid = blockIdx.x * (size_t)256 + threadIdx.x;
    const bool vec = (((uintptr_t)s | (uintptr_t)m) % 16) == 0;
    size_t nvec = vec ? npix / 16 : 0;
    // per-byte compare: (x > t) -> 0xff.  __vcmpgtu4 has no native SASS; use the
    // carry trick: ((x | 0x100) - (t + 1)) bit 8 set iff x > t, per 16-bit lane.
    for (size_t i = gid; i < nvec; i += stride) {
        uint4 v = ld_stream(reinterpret_cast<const uint4*>(s) + i);
        uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t x = wds[j];
            uint32_t lo = x & 0x00ff00ffu, hi = (x >> 8) & 0x00ff00ffu;
            uint32_t tt = (t + 1u) * 0x00010001u;
            uint32_t clo = (((lo | 0x01000100u) - tt) >> 8) & 0x00010001u;
            uint32_t chi = (((hi | 0x01000100u) - tt) >> 8) & 0x00010001u;
            wds[j] = ((clo * 255u) | ((chi * 255u) << 8)) ^ flip;
        }
        reinterpret_cast<uint4*>(m)[i] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    }
    for (size_t i = nvec * 16 + gid; i < npix; i += stride) m[i] = (uint8_t)(((s[i] > t) ? 255u : 0u) ^ (flip & 255u));
}

}  // namespace

int launch_adaptive(llfe_ctx* ctx, const uint8_t* gray, int n, int h, int w, int C, uint8_t* mask, uint64_t* sum_count) {
    if (n == 0 || h == 0 || w == 0) return LLFE_OK;
    if (sum_count) LLFE_CUDA(cudaMemsetAsync(sum_count, 0, (size_t)n * 2 * sizeof(uint64_t), ctx->stream));
    dim3 grid(ceil_div(w, ATW), ceil_div(h, ATH), n);
    LLFE_KERNEL(ctx, "k_adaptive");
    k_adaptive<<<grid, 256, 0, ctx->stream>>>(gray, h, w, C, mask, (unsigned long long*)sum_count);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

int launch_hist256(llfe_ctx* ctx, const uint8_t* gray, int n, size_t npix, uint32_t* hist) {
    LLFE_CUDA(cudaMemsetAsync(hist, 0, (size_t)n * 256 * sizeof(uint32_t), ctx->stream));
    if (n == 0 || npix == 0) return LLFE_OK;
    size_t want = ceil_div_sz(npix, 16 * 256 * 4);
    unsigned gx = (unsigned)(want < 1 ? 1 : (want > 1024 ? 1024 : want));
    LLFE_KERNEL(ctx, "k_hist256");
    k_hist256<<<dim3(gx, n), 256, 0, ctx->stream>>>(gray, npix, hist);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

int launch_otsu_sweep(llfe_ctx* ctx, const uint32_t* hist, int n, size_t npix, int invert_if_light, int32_t* thresh,
                      int32_t* invert) {
    if (n == 0) return LLFE_OK;
    LLFE_KERNEL(ctx, "k_otsu_sweep");
    k_otsu_sweep<<<ceil_div(n, 64), 64, 0, ctx->stream>>>(hist, n, npix, invert_if_light, thresh, invert);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

int launch_binarize(llfe_ctx* ctx, const uint8_t* gray, int n, size_t npix, const int32_t* thresh, const int32_t* invert,
                    uint8_t* mask) {
    if (n == 0 || npix == 0) return LLFE_OK;
    size_t want = ceil_div_sz(npix, 16 * 256 * 2);
    unsigned gx = (unsigned)(want < 1 ? 1 : (want > 2048 ? 2048 : want));
    LLFE_KERNEL(ctx, "k_binarize");
    k_binarize<<<dim3(gx, n), 256, 0, ctx->stream>>>(gray, npix, thresh, invert, mask);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}
