// Pillow's LANCZOS down-scaler on u8 images: the arithmetic behind
//     pil_image.thumbnail((max_width, max_height), Image.Resampling.LANCZOS)
// in the reference's ImageProcessor.auto_process_image (app/services/analyze/image_processor.py:221-224).
// `Image.resize(size, LANCZOS, box, reducing_gap=2.0)` is an optional integer box reduction (`ImagingReduce`) followed by
// the two-pass resampler (`ImagingResample`); the Python layer (services/image_processor.py) mirrors PIL/Image.py's size,
// factor and box rules and calls the two primitives below.  Restated in oracle/pilops.py and pinned against the installed
// Pillow binary (tests/test_oracle_pil.py).
//
//   k_pil_reduce     out = ((sum of the fx x fy cell + n / 2) * floor(2^24 / n)) >> 24, n = pixels under the cell (cells at
//                    the right / bottom edge of the box are partial and use their own n).
//   k_pil_resample_h / _v
//                    per output index a window [first, first + count) and 22-bit fixed-point weights
//                    (double-precision lanczos((x + first - center + 0.5) / filterscale), normalised, rounded half away
//                    from zero -- built on the host with libm like Pillow does, cached per context); horizontal pass
//                    over the rows the vertical pass needs into a u8 intermediate, then the vertical pass;
//                    out = clip8((2^21 + sum pixel * weight) >> 22).
#include <math.h>

#include <vector>

#include "llfe_common.cuh"
#include "llfe_device.cuh"

namespace {

constexpr int PIL_PRECISION_BITS = 32 - 8 - 2;

struct PilTab {
    int ksize = 0;
    std::vector<int> first, count, kk;   // kk is [tap][out index]
};

double pil_sinc(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return sin(x) / x;
}

double pil_lanczos(double x) { return (-3.0 <= x && x < 3.0) ? pil_sinc(x) * pil_sinc(x / 3) : 0.0; }

void pil_coeffs(int in_size, float in0, float in1, int out_size, PilTab* t) {
    const double scale = (double)(in1 - in0) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 3.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    t->ksize = ksize;
    t->first.assign(out_size, 0);
    t->count.assign(out_size, 0);
    t->kk.assign((size_t)ksize * out_size, 0);
    std::vector<double> k(ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = in0 + (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            const double w = pil_lanczos((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x) {
            const double v = ww != 0.0 ? k[x] / ww : k[x];
            t->kk[(size_t)x * out_size + xx] =
                v < 0 ? (int)(-0.5 + v * (1 << PIL_PRECISION_BITS)) : (int)(0.5 + v * (1 << PIL_PRECISION_BITS));
        }
        t->first[xx] = xmin;
        t->count[xx] = xmax;
    }
}

__device__ __forceinline__ uint8_t pil_clip8(int acc) {
    const int v = acc >> PIL_PRECISION_BITS;
    return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
}

// src (n, sh, sw, C) rows [row0, row0 + rows) -> tmp (n, rows, dw, C)
template <int C>
__global__ void __launch_bounds__(256) k_pil_resample_h(const uint8_t* __restrict__ src, int sh, int sw, int row0, int rows,
                                                        uint8_t* __restrict__ tmp, int dw, const int* __restrict__ first,
                                                        const int* __restrict__ count, const int* __restrict__ kk) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= dw) return;
    const uint8_t* p = src + (((size_t)img * sh + row0 + y) * sw + first[x]) * C;
    const int n = count[x];
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (PIL_PRECISION_BITS - 1);
    for (int k = 0; k < n; ++k) {
        const int w = kk[(size_t)k * dw + x];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += (int)p[k * C + c] * w;
    }
    uint8_t* o = tmp + (((size_t)img * rows + y) * dw + x) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = pil_clip8(acc[c]);
}

// src (n, rows, wc) u8 (wc = width * channels) -> dst (n, dh, wc); first[] is relative to row `shift` of src
__global__ void __launch_bounds__(256) k_pil_resample_v(const uint8_t* __restrict__ src, int rows, int wc, int shift,
                                                        uint8_t* __restrict__ dst, int dh, const int* __restrict__ first,
                                                        const int* __restrict__ count, const int* __restrict__ kk) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= wc) return;
    const uint8_t* p = src + ((size_t)img * rows + first[y] - shift) * wc + x;
    const int n = count[y];
    int acc = 1 << (PIL_PRECISION_BITS - 1);
    for (int k = 0; k < n; ++k) acc += (int)p[(size_t)k * wc] * kk[(size_t)k * dh + y];
    dst[((size_t)img * dh + y) * wc + x] = pil_clip8(acc);
}

template <int C>
__global__ void __launch_bounds__(256) k_pil_reduce(const uint8_t* __restrict__ src, int sh, int sw, int bx0, int by0, int bw,
                                                    int bh, int fx, int fy, uint8_t* __restrict__ dst, int dh, int dw) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= dw) return;
    const int nx = min(fx, bw - x * fx), ny = min(fy, bh - y * fy);
    const uint8_t* p = src + (((size_t)img * sh + by0 + y * fy) * sw + bx0 + x * fx) * C;
    uint32_t s[C];
#pragma unroll
    for (int c = 0; c < C; ++c) s[c] = 0;
    for (int yy = 0; yy < ny; ++yy)
        for (int xx = 0; xx < nx; ++xx) {
#pragma unroll
            for (int c = 0; c < C; ++c) s[c] += p[((size_t)yy * sw + xx) * C + c];
        }
    const uint32_t n = (uint32_t)(nx * ny), mul = (1u << 24) / n;
    uint8_t* o = dst + (((size_t)img * dh + y) * dw + x) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = (uint8_t)(((unsigned long long)(s[c] + n / 2) * mul) >> 24);
}

}  // namespace

extern "C" int llfe_pil_reduce(llfe_ctx* ctx, const uint8_t* d_src, int n, int sh, int sw, int c, const int32_t* box, int fx,
                               int fy, uint8_t* d_dst) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_src != nullptr && d_dst != nullptr && box != nullptr && (c == 1 || c == 3));
    LLFE_CHECK_ARG(n >= 0 && n <= 65535 && sh > 0 && sw > 0 && fx >= 1 && fy >= 1 && (int64_t)fx * fy <= 65536);
    LLFE_CHECK_ARG(box[0] >= 0 && box[1] >= 0 && box[2] > box[0] && box[3] > box[1] && box[2] <= sw && box[3] <= sh);
    if (n == 0) return LLFE_OK;
    const int bw = box[2] - box[0], bh = box[3] - box[1];
    const int dw = ceil_div(bw, fx), dh = ceil_div(bh, fy);
    LLFE_CHECK_ARG(dh <= 65535);
    dim3 grid(ceil_div(dw, 256), dh, n);
    LLFE_KERNEL(ctx, "k_pil_reduce");
    if (c == 3)
        k_pil_reduce<3><<<grid, 256, 0, ctx->stream>>>(d_src, sh, sw, box[0], box[1], bw, bh, fx, fy, d_dst, dh, dw);
    else
        k_pil_reduce<1><<<grid, 256, 0, ctx->stream>>>(d_src, sh, sw, box[0], box[1], bw, bh, fx, fy, d_dst, dh, dw);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

extern "C" int llfe_pil_resample_lanczos(llfe_ctx* ctx, const uint8_t* d_src, int n, int sh, int sw, int c, const float* box,
                                         uint8_t* d_dst, int dh, int dw) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_src != nullptr && d_dst != nullptr && box != nullptr && (c == 1 || c == 3));
    LLFE_CHECK_ARG(n >= 0 && n <= 65535 && sh > 0 && sw > 0 && dh > 0 && dw > 0 && dh <= 65535 && sh <= 65535);
    LLFE_CHECK_ARG(box[0] >= 0.f && box[1] >= 0.f && box[2] > box[0] && box[3] > box[1] && box[2] <= (float)sw &&
                   box[3] <= (float)sh);
    if (n == 0) return LLFE_OK;
    const bool need_h = dw != sw || box[0] != 0.f || box[2] != (float)dw;
    const bool need_v = dh != sh || box[1] != 0.f || box[3] != (float)dh;
    if (!need_h && !need_v) {
        LLFE_CUDA(cudaMemcpyAsync(d_dst, d_src, (size_t)n * sh * sw * c, cudaMemcpyDeviceToDevice, ctx->stream));
        return LLFE_OK;
    }
    PilTab th, tv;
    pil_coeffs(sw, box[0], box[2], dw, &th);
    pil_coeffs(sh, box[1], box[3], dh, &tv);
    const int row_first = tv.first[0], row_last = tv.first[dh - 1] + tv.count[dh - 1];
    const int rows = row_last - row_first;
    // tables + the horizontally resampled rows in the workspace
    const size_t tab_ints = 2 * (size_t)dw + th.kk.size() + 2 * (size_t)dh + tv.kk.size();
    const size_t tmp_bytes = need_h && need_v ? (size_t)n * rows * dw * c : 0;
    void* ws;
    LLFE_TRY(llfe_workspace(ctx, WsCarver::need(tab_ints * 4 + 6 * 256) + WsCarver::need(tmp_bytes), &ws));
    WsCarver carve(ws);
    int* d_hf = carve.take<int>(dw);
    int* d_hc = carve.take<int>(dw);
    int* d_hk = carve.take<int>(th.kk.size());
    int* d_vf = carve.take<int>(dh);
    int* d_vc = carve.take<int>(dh);
    int* d_vk = carve.take<int>(tv.kk.size());
    uint8_t* tmp = carve.take<uint8_t>(tmp_bytes);
    // pageable sources: the copies are staged by the runtime before the call returns
    LLFE_CUDA(cudaMemcpyAsync(d_hf, th.first.data(), dw * 4, cudaMemcpyHostToDevice, ctx->stream));
    LLFE_CUDA(cudaMemcpyAsync(d_hc, th.count.data(), dw * 4, cudaMemcpyHostToDevice, ctx->stream));
    LLFE_CUDA(cudaMemcpyAsync(d_hk, th.kk.data(), th.kk.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    LLFE_CUDA(cudaMemcpyAsync(d_vf, tv.first.data(), dh * 4, cudaMemcpyHostToDevice, ctx->stream));
    LLFE_CUDA(cudaMemcpyAsync(d_vc, tv.count.data(), dh * 4, cudaMemcpyHostToDevice, ctx->stream));
    LLFE_CUDA(cudaMemcpyAsync(d_vk, tv.kk.data(), tv.kk.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    const uint8_t* vsrc = d_src;
    int vrows = sh, vshift = 0;
    if (need_h) {
        // without a vertical pass every row is an output row
        const int r0 = need_v ? row_first : 0, nr = need_v ? rows : sh;
        uint8_t* out = need_v ? tmp : d_dst;
        dim3 grid(ceil_div(dw, 256), nr, n);
        LLFE_KERNEL(ctx, "k_pil_resample_h");
        if (c == 3)
            k_pil_resample_h<3><<<grid, 256, 0, ctx->stream>>>(d_src, sh, sw, r0, nr, out, dw, d_hf, d_hc, d_hk);
        else
            k_pil_resample_h<1><<<grid, 256, 0, ctx->stream>>>(d_src, sh, sw, r0, nr, out, dw, d_hf, d_hc, d_hk);
        LLFE_LAUNCHED(ctx);
        vsrc = tmp;
        vrows = rows;
        vshift = row_first;
    }
    if (need_v) {
        const int wc = dw * c;
        dim3 grid(ceil_div(wc, 256), dh, n);
        LLFE_KERNEL(ctx, "k_pil_resample_v");
        k_pil_resample_v<<<grid, 256, 0, ctx->stream>>>(vsrc, vrows, wc, vshift, d_dst, dh, d_vf, d_vc, d_vk);
        LLFE_LAUNCHED(ctx);
    }
    // the host tables go out of scope on return: wait for the staged copies
    LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
    return LLFE_OK;
}
