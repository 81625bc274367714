// cv2.resize(..., interpolation=INTER_AREA) for down-scaling u8 images (1 or 3
// channels), all three OpenCV code paths:
//   * both ratios integer and == 2 : (sum of 2x2 + 2) >> 2
//   * both ratios integer          : sat_u8(rint(float(sum) * (1.f / area)))
//   * otherwise                    : per-axis (index, float32 weight) tables; the
//     horizontal pass accumulates  buf = f32(buf + f32(S * alpha))  in table order
//     (separate multiply and add), the vertical pass  sum = beta * buf  for the
//     first entry and  sum = f32(sum + f32(beta * buf))  afterwards.
// The tables are built on the host in float64 exactly as OpenCV builds them and
// cached per (source size, destination size) on the context.
#include <math.h>

#include <vector>

#include "llfe_common.cuh"
#include "llfe_device.cuh"

struct AreaTab {
    int kind;  // 0: INTER_AREA (offsets / indices / float weights); 1: LANCZOS4 (d_idx = first tap - 3, d_ofs = 8 weights)
    int ssize, dsize;
    int n_entries;
    int max_per_dst;
    int* d_ofs;  // dsize + 1 offsets into the entry arrays
    int* d_idx;
    float* d_w;
    AreaTab* next;
};

namespace {

void build_tab(int ssize, int dsize, double scale, std::vector<int>& ofs, std::vector<int>& idx, std::vector<float>& wt) {
    ofs.assign(dsize + 1, 0);
    for (int d = 0; d < dsize; ++d) {
        ofs[d] = (int)idx.size();
        double fsx1 = d * scale, fsx2 = fsx1 + scale;
        double cell = fmin(scale, ssize - fsx1);
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
        sx1 = sx1 < sx2 ? sx1 : sx2;
        if (sx1 - fsx1 > 1e-3) {
            idx.push_back(sx1 - 1);
            wt.push_back((float)((sx1 - fsx1) / cell));
        }
        for (int sx = sx1; sx < sx2; ++sx) {
            idx.push_back(sx);
            wt.push_back((float)(1.0 / cell));
        }
        if (fsx2 - sx2 > 1e-3) {
            idx.push_back(sx2);
            wt.push_back((float)(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell));
        }
    }
    ofs[dsize] = (int)idx.size();
}

int get_tab(llfe_ctx* ctx, int ssize, int dsize, AreaTab** out) {
    for (AreaTab* t = ctx->area_tabs; t; t = t->next)
        if (t->kind == 0 && t->ssize == ssize && t->dsize == dsize) {
            *out = t;
            return LLFE_OK;
        }
    std::vector<int> ofs, idx;
    std::vector<float> wt;
    const double scale = 1.0 / ((double)dsize / (double)ssize);
    build_tab(ssize, dsize, scale, ofs, idx, wt);
    AreaTab* t = new AreaTab();
    t->kind = 0;
    t->ssize = ssize;
    t->dsize = dsize;
    t->n_entries = (int)idx.size();
    t->max_per_dst = 0;
    for (int d = 0; d < dsize; ++d) t->max_per_dst = ofs[d + 1] - ofs[d] > t->max_per_dst ? ofs[d + 1] - ofs[d] : t->max_per_dst;
    t->d_ofs = nullptr;
    t->d_idx = nullptr;
    t->d_w = nullptr;
    t->next = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&t->d_ofs, ofs.size() * sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc(&t->d_idx, (idx.size() + 1) * sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc(&t->d_w, (wt.size() + 1) * sizeof(float))) != cudaSuccess) {
        if (t->d_ofs) cudaFree(t->d_ofs);
        if (t->d_idx) cudaFree(t->d_idx);
        delete t;
        return llfe_cuda_fail(e, "cudaMalloc(area table)", __FILE__, __LINE__);
    }
    // synchronous copies: the host vectors die at the end of this function
    LLFE_CUDA(cudaMemcpy(t->d_ofs, ofs.data(), ofs.size() * sizeof(int), cudaMemcpyHostToDevice));
    LLFE_CUDA(cudaMemcpy(t->d_idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
    LLFE_CUDA(cudaMemcpy(t->d_w, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
    t->next = ctx->area_tabs;
    ctx->area_tabs = t;
    *out = t;
    return LLFE_OK;
}

template <int C>
__global__ void __launch_bounds__(256) k_resize_area_tab(const uint8_t* __restrict__ src, int sh, int sw,
                                                         uint8_t* __restrict__ dst, int dh, int dw,
                                                         const int* __restrict__ xofs, const int* __restrict__ xidx,
                                                         const float* __restrict__ xw, const int* __restrict__ yofs,
                                                         const int* __restrict__ yidx, const float* __restrict__ yw) {
    const int img = blockIdx.z;
    const uint8_t* s = src + (size_t)img * sh * sw * C;
    uint8_t* d = dst + (size_t)img * dh * dw * C;
    int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= dw) return;
    const int xb = xofs[dx], xe = xofs[dx + 1];
    const int yb = yofs[dy], ye = yofs[dy + 1];
    float sum[C];
    for (int k = yb; k < ye; ++k) {
        const uint8_t* row = s + (size_t)yidx[k] * sw * C;
        const float beta = yw[k];
        float buf[C];
#pragma unroll
        for (int ch = 0; ch < C; ++ch) buf[ch] = 0.f;
        for (int j = xb; j < xe; ++j) {
            const uint8_t* p = row + (size_t)xidx[j] * C;
            const float alpha = xw[j];
#pragma unroll
            for (int ch = 0; ch < C; ++ch) buf[ch] = __fadd_rn(buf[ch], __fmul_rn((float)p[ch], alpha));
        }
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            float t = __fmul_rn(beta, buf[ch]);
            sum[ch] = (k == yb) ? t : __fadd_rn(sum[ch], t);
        }
    }
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        int v = __float2int_rn(sum[ch]);
        d[((size_t)dy * dw + dx) * C + ch] = (uint8_t)min(max(v, 0), 255);
    }
}

template <int C>
__global__ void __launch_bounds__(256) k_resize_area_int(const uint8_t* __restrict__ src, int sh, int sw,
                                                         uint8_t* __restrict__ dst, int dh, int dw, int isx, int isy) {
    const int img = blockIdx.z;
    const uint8_t* s = src + (size_t)img * sh * sw * C;
    uint8_t* d = dst + (size_t)img * dh * dw * C;
    int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= dw) return;
    int sum[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) sum[ch] = 0;
    for (int yy = 0; yy < isy; ++yy) {
        const uint8_t* row = s + ((size_t)(dy * isy + yy) * sw + (size_t)dx * isx) * C;
        for (int xx = 0; xx < isx; ++xx)
#pragma unroll
            for (int ch = 0; ch < C; ++ch) sum[ch] += row[xx * C + ch];
    }
    const bool two = (isx == 2 && isy == 2);
    const float inv = __fdiv_rn(1.f, (float)(isx * isy));
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        int v = two ? ((sum[ch] + 2) >> 2) : __float2int_rn(__fmul_rn((float)sum[ch], inv));
        d[((size_t)dy * dw + dx) * C + ch] = (uint8_t)min(max(v, 0), 255);
    }
}

}  // namespace

void llfe_free_area_tabs(llfe_ctx* ctx) {
    AreaTab* t = ctx->area_tabs;
    while (t) {
        AreaTab* nx = t->next;
        cudaFree(t->d_ofs);
        cudaFree(t->d_idx);
        if (t->d_w) cudaFree(t->d_w);
        delete t;
        t = nx;
    }
    ctx->area_tabs = nullptr;
}

extern "C" int llfe_resize_area(llfe_ctx* ctx, const uint8_t* d_src, int n, int sh, int sw, int c, uint8_t* d_dst,
                                int dh, int dw) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_src != nullptr && d_dst != nullptr);
    LLFE_CHECK_ARG(n >= 0 && n <= 65535 && sh > 0 && sw > 0 && dh > 0 && dw > 0 && (c == 1 || c == 3));
    if (dh > sh || dw > sw) {
        llfe_set_error("llfe_resize_area: only down-scaling is on the hot path (%dx%d -> %dx%d)", sw, sh, dw, dh);
        return LLFE_E_UNSUPPORTED;
    }
    if (dh > 65535) return LLFE_E_UNSUPPORTED;
    if (n == 0) return LLFE_OK;
    if (dh == sh && dw == sw) {
        LLFE_CUDA(cudaMemcpyAsync(d_dst, d_src, (size_t)n * sh * sw * c, cudaMemcpyDeviceToDevice, ctx->stream));
        return LLFE_OK;
    }
    const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    const int isx = (int)lrint(scale_x), isy = (int)lrint(scale_y);
    const bool fast = fabs(scale_x - isx) < 2.220446049250313e-16 && fabs(scale_y - isy) < 2.220446049250313e-16;
    dim3 grid(ceil_div(dw, 256), dh, n);
    if (fast) {
        LLFE_KERNEL(ctx, "k_resize_area_int");
        if (c == 3)
            k_resize_area_int<3><<<grid, 256, 0, ctx->stream>>>(d_src, sh, sw, d_dst, dh, dw, isx, isy);
        else
            k_resize_area_int<1><<<grid, 256, 0, ctx->stream>>>(d_src, sh, sw, d_dst, dh, dw, isx, isy);
        LLFE_LAUNCHED(ctx);
        return LLFE_OK;
    }
    AreaTab *xt, *yt;
    LLFE_TRY(get_tab(ctx, sw, dw, &xt));
    LLFE_TRY(get_tab(ctx, sh, dh, &yt));
    LLFE_KERNEL(ctx, "k_resize_area_tab");
    if (c == 3)
        k_resize_area_tab<3><<<grid, 256, 0, ctx->stream>>>(d_src, sh, sw, d_dst, dh, dw, xt->d_ofs, xt->d_idx, xt->d_w,
                                                            yt->d_ofs, yt->d_idx, yt->d_w);
    else
        k_resize_area_tab<1><<<grid, 256, 0, ctx->stream>>>(d_src, sh, sw, d_dst, dh, dw, xt->d_ofs, xt->d_idx, xt->d_w,
                                                            yt->d_ofs, yt->d_idx, yt->d_w);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

// ---------------------------------------------------------------------------------------------------
// cv2.resize(..., interpolation=INTER_LINEAR) on u8 (the `performance` preprocessing mode,
// app/services/analyze/utils.py:136-143).  OpenCV's fixed-point path, restated (oracle/cvops.py
// resize_linear, verified against cv2 bit for bit):
//   per axis  f = float((d + 0.5) * scale - 0.5) (double arithmetic), s = floor(f), f -= s;
//             x only: s < 0 -> (0, f = 0);  s >= ssize - 1 -> (ssize - 1, f = 0);
//             y: the fraction is kept and the two row indices are clipped into the image;
//             weights = short(rint((1 - f) * 2048)), short(rint(f * 2048))
//   H(row)    = S[row][x0] * a0 + S[row][x1] * a1                      (int32)
//   dst       = sat_u8((((b0 * (H(y0) >> 4)) >> 16) + ((b1 * (H(y1) >> 4)) >> 16) + 2) >> 2)
// and when both ratios are exactly 2 OpenCV takes the INTER_AREA 2x2 path instead.
namespace {

struct LinCoef {
    int s0, s1;
    int w0, w1;
};

template <bool CLAMP_WEIGHTS>
__device__ __forceinline__ LinCoef lin_coef(int d, double scale, int ssize) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (CLAMP_WEIGHTS) {
        if (s < 0) {
            f = 0.f;
            s = 0;
        }
        if (s >= ssize - 1) {
            f = 0.f;
            s = ssize - 1;
        }
    }
    LinCoef c;
    c.s0 = min(max(s, 0), ssize - 1);
    c.s1 = min(max(s + 1, 0), ssize - 1);
    c.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    c.w1 = __float2int_rn(__fmul_rn(f, 2048.f));
    return c;
}

template <int C>
__global__ void __launch_bounds__(256) k_resize_linear(const uint8_t* __restrict__ src, int sh, int sw,
                                                       uint8_t* __restrict__ dst, int dh, int dw, double scale_x,
                                                       double scale_y) {
    const int img = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= dw) return;
    const LinCoef cx = lin_coef<true>(x, scale_x, sw), cy = lin_coef<false>(y, scale_y, sh);
    const uint8_t* s = src + (size_t)img * sh * sw * C;
    const uint8_t* r0 = s + (size_t)cy.s0 * sw * C;
    const uint8_t* r1 = s + (size_t)cy.s1 * sw * C;
    uint8_t* o = dst + (((size_t)img * dh + y) * dw + x) * C;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        const int h0 = (int)r0[cx.s0 * C + ch] * cx.w0 + (int)r0[cx.s1 * C + ch] * cx.w1;
        const int h1 = (int)r1[cx.s0 * C + ch] * cx.w0 + (int)r1[cx.s1 * C + ch] * cx.w1;
        const int v = (((cy.w0 * (h0 >> 4)) >> 16) + ((cy.w1 * (h1 >> 4)) >> 16) + 2) >> 2;
        o[ch] = (uint8_t)min(max(v, 0), 255);
    }
}

}  // namespace

extern "C" int llfe_resize_linear(llfe_ctx* ctx, const uint8_t* d_src, int n, int sh, int sw, int c, uint8_t* d_dst,
                                  int dh, int dw) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_src != nullptr && d_dst != nullptr);
    LLFE_CHECK_ARG(n >= 0 && n <= 65535 && sh > 0 && sw > 0 && dh > 0 && dw > 0 && dh <= 65535 && (c == 1 || c == 3));
    if (n == 0) return LLFE_OK;
    if (sw == 2 * dw && sh == 2 * dh) return llfe_resize_area(ctx, d_src, n, sh, sw, c, d_dst, dh, dw);  // OpenCV does the same
    const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    dim3 grid(ceil_div(dw, 256), dh, n);
    LLFE_KERNEL(ctx, "k_resize_linear");
    if (c == 3)
        k_resize_linear<3><<<grid, 256, 0, ctx->stream>>>(d_src, sh, sw, d_dst, dh, dw, scale_x, scale_y);
    else
        k_resize_linear<1><<<grid, 256, 0, ctx->stream>>>(d_src, sh, sw, d_dst, dh, dw, scale_x, scale_y);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

// ---------------------------------------------------------------------------------------------------
// cv2.resize(..., interpolation=INTER_LANCZOS4) on u8 (the `high_quality` preprocessing mode,
// app/services/analyze/utils.py:128-135).  OpenCV's fixed-point path, restated (oracle/cvops.py
// resize_lanczos4, verified against cv2 bit for bit):
//   per axis  f = float((d + 0.5) * scale - 0.5), s = floor(f), f -= s; 8 float coefficients from
//             interpolateLanczos4(f) (sin / cos of the first tap in double, rotated by 45 degrees per tap,
//             divided by y^2, normalised in float), each scaled by 2048 and rounded to short;
//             taps s - 3 .. s + 4, indices clipped into the image (no weight clamping on either axis);
//   H(row)  = sum_k S[row][x_k] * a_k            (int)
//   dst     = sat_u8((sum_k H(y_k) * b_k + 2^21) >> 22)
// The tables are built on the host with the same libm calls as OpenCV's and cached on the context.
namespace {

void lanczos4_coeffs(float x, float* coeffs) {
    static const double s45 = 0.70710678118654752440084436210485;
    static const double cs[][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    const double kPi = 3.1415926535897932384626433832795;
    float sum = 0;
    const double y0 = -(x + 3) * kPi * 0.25, s0 = sin(y0), c0 = cos(y0);
    for (int i = 0; i < 8; i++) {
        const float y0_ = (x + 3 - i);
        if (fabs(y0_) >= 1e-6f) {
            const double y = -y0_ * kPi * 0.25;
            coeffs[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
        } else {
            coeffs[i] = 1e30f;  // x ~ 0 or ~ 1: this tap takes everything after the normalisation
        }
        sum += coeffs[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < 8; i++) coeffs[i] *= sum;
}

int get_lanczos_tab(llfe_ctx* ctx, int ssize, int dsize, AreaTab** out) {
    for (AreaTab* t = ctx->area_tabs; t; t = t->next)
        if (t->kind == 1 && t->ssize == ssize && t->dsize == dsize) {
            *out = t;
            return LLFE_OK;
        }
    const double scale = 1.0 / ((double)dsize / (double)ssize);
    std::vector<int> first(dsize), wts((size_t)dsize * 8);
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        const int sx = (int)floorf(f);
        f -= sx;
        float c[8];
        lanczos4_coeffs(f, c);
        first[d] = sx - 3;
        for (int k = 0; k < 8; ++k) {
            long v = lrintf(c[k] * 2048.f);   // saturate_cast<short>(float): round half to even, then clamp
            wts[(size_t)d * 8 + k] = (int)(v < -32768 ? -32768 : v > 32767 ? 32767 : v);
        }
    }
    AreaTab* t = new AreaTab();
    t->kind = 1;
    t->ssize = ssize;
    t->dsize = dsize;
    t->n_entries = dsize * 8;
    t->max_per_dst = 8;
    t->d_ofs = nullptr;
    t->d_idx = nullptr;
    t->d_w = nullptr;
    t->next = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&t->d_ofs, wts.size() * sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc(&t->d_idx, first.size() * sizeof(int))) != cudaSuccess) {
        if (t->d_ofs) cudaFree(t->d_ofs);
        delete t;
        return llfe_cuda_fail(e, "cudaMalloc(lanczos table)", __FILE__, __LINE__);
    }
    LLFE_CUDA(cudaMemcpy(t->d_ofs, wts.data(), wts.size() * sizeof(int), cudaMemcpyHostToDevice));
    LLFE_CUDA(cudaMemcpy(t->d_idx, first.data(), first.size() * sizeof(int), cudaMemcpyHostToDevice));
    t->next = ctx->area_tabs;
    ctx->area_tabs = t;
    *out = t;
    return LLFE_OK;
}

template <int C>
__global__ void __launch_bounds__(256) k_resize_lanczos4(const uint8_t* __restrict__ src, int sh, int sw,
                                                         uint8_t* __restrict__ dst, int dh, int dw,
                                                         const int* __restrict__ x_first, const int* __restrict__ x_w,
                                                         const int* __restrict__ y_first, const int* __restrict__ y_w) {
    const int img = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= dw) return;
    const uint8_t* s = src + (size_t)img * sh * sw * C;
    int xs[8], xa[8];
    const int x0 = x_first[x], y0 = y_first[y];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        xs[k] = min(max(x0 + k, 0), sw - 1) * C;
        xa[k] = x_w[x * 8 + k];
    }
    long long acc[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) acc[ch] = 0;
#pragma unroll
    for (int ky = 0; ky < 8; ++ky) {
        const uint8_t* row = s + (size_t)min(max(y0 + ky, 0), sh - 1) * sw * C;
        const int b = y_w[y * 8 + ky];
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            int hsum = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) hsum += (int)row[xs[k] + ch] * xa[k];
            acc[ch] += (long long)hsum * b;
        }
    }
    uint8_t* o = dst + (((size_t)img * dh + y) * dw + x) * C;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        const long long v = (acc[ch] + (1ll << 21)) >> 22;
        o[ch] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
}

}  // namespace

extern "C" int llfe_resize_lanczos4(llfe_ctx* ctx, const uint8_t* d_src, int n, int sh, int sw, int c, uint8_t* d_dst,
                                    int dh, int dw) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_src != nullptr && d_dst != nullptr);
    LLFE_CHECK_ARG(n >= 0 && n <= 65535 && sh > 0 && sw > 0 && dh > 0 && dw > 0 && dh <= 65535 && (c == 1 || c == 3));
    if (n == 0) return LLFE_OK;
    AreaTab *xt, *yt;
    LLFE_TRY(get_lanczos_tab(ctx, sw, dw, &xt));
    LLFE_TRY(get_lanczos_tab(ctx, sh, dh, &yt));
    dim3 grid(ceil_div(dw, 256), dh, n);
    LLFE_KERNEL(ctx, "k_resize_lanczos4");
    if (c == 3)
        k_resize_lanczos4<3><<<grid, 256, 0, ctx->stream>>>(d_src, sh, sw, d_dst, dh, dw, xt->d_idx, xt->d_ofs, yt->d_idx,
                                                            yt->d_ofs);
    else
        k_resize_lanczos4<1><<<grid, 256, 0, ctx->stream>>>(d_src, sh, sw, d_dst, dh, dw, xt->d_idx, xt->d_ofs, yt->d_idx,
                                                            yt->d_ofs);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}
