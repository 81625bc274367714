// 5x5 binomial Gaussian blur on u8 (cv2.GaussianBlur(src,(5,5),0)):
// separable [1,4,6,4,1] both ways, BORDER_REFLECT_101, (sum + 128) >> 8.
// One kernel, two front ends: generic C-channel input, or BGR input converted to
// gray on the fly (the fused gray+blur of ShadowAnalyzer.preprocess_image).
#include "llfe_common.cuh"
#include "llfe_device.cuh"

namespace {

constexpr int TWB = 128;  // tile width in output bytes
constexpr int TH = 32;    // tile height in rows
constexpr int MAXC = 3;

// FROM_BGR: src is (h, w, 3) BGR, output is 1-channel gray-blurred.
// otherwise : src is (h, w, C), output (h, w, C), blur per channel.
template <bool FROM_BGR>
__global__ void __launch_bounds__(256) k_blur5(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int w,
                                               int c) {
    __shared__ uint8_t tin[TH + 4][TWB + 4 * MAXC + 4];
    __shared__ uint16_t hs[TH + 4][TWB];
    const int C = FROM_BGR ? 1 : c;  // channels of the blurred signal
    const int wb = w * C;            // output row bytes
    const int img = blockIdx.z;
    const size_t in_img = (size_t)h * w * (FROM_BGR ? 3 : C);
    const uint8_t* s = src + img * in_img;
    uint8_t* d = dst + (size_t)img * h * wb;
    const int xb0 = blockIdx.x * TWB;  // first output byte of the tile
    const int y0 = blockIdx.y * TH;
    const int tid = threadIdx.x;
    const int halo = 2 * C;
    const int inw = TWB + 2 * halo;

    // phase 1: stage the (TH+4) x (TWB + 4C) input bytes with reflect-101 on pixel indices
    for (int i = tid; i < (TH + 4) * inw; i += 256) {
        int ry = i / inw, rx = i - ry * inw;
        int y = reflect101(y0 - 2 + ry, h);
        int xb = xb0 - halo + rx;  // byte position in the (virtual) output row
        int px = (xb >= 0) ? xb / C : -((-xb + C - 1) / C);
        int ch = xb - px * C;
        px = reflect101(px, w);
        uint8_t v;
        if (FROM_BGR) {
            const uint8_t* p = s + ((size_t)y * w + px) * 3;
            v = gray_px(p[0], p[1], p[2]);
        } else {
            v = s[((size_t)y * w + px) * C + ch];
        }
        tin[ry][rx] = v;
    }
    __syncthreads();
    // phase 2: horizontal pass
    for (int i = tid; i < (TH + 4) * TWB; i += 256) {
        int ry = i / TWB, rx = i - ry * TWB;
        const uint8_t* p = &tin[ry][rx + halo];
        hs[ry][rx] = (uint16_t)(p[-2 * C] + 4 * p[-C] + 6 * p[0] + 4 * p[C] + p[2 * C]);
    }
    __syncthreads();
    // phase 3: vertical pass + store (4 bytes per thread-step)
    for (int i = tid; i < TH * (TWB / 4); i += 256) {
        int ry = i / (TWB / 4), rx = (i - ry * (TWB / 4)) * 4;
        int y = y0 + ry;
        if (y >= h) continue;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v = hs[ry][rx + k] + 4u * hs[ry + 1][rx + k] + 6u * hs[ry + 2][rx + k] + 4u * hs[ry + 3][rx + k] +
                         hs[ry + 4][rx + k];
            out |= ((v + 128u) >> 8) << (8 * k);
        }
        int xb = xb0 + rx;
        uint8_t* o = d + (size_t)y * wb + xb;
        if (xb + 3 < wb && (((uintptr_t)o) & 3) == 0) {
            *reinterpret_cast<uint32_t*>(o) = out;
        } else {
            for (int k = 0; k < 4; ++k)
                if (xb + k < wb) o[k] = (uint8_t)(out >> (8 * k));
        }
    }
}

}  // namespace

int launch_blur5(llfe_ctx* ctx, const uint8_t* src, int n, int h, int w, int c, uint8_t* dst) {
    if (n == 0 || h == 0 || w == 0) return LLFE_OK;
    dim3 grid(ceil_div(w * c, TWB), ceil_div(h, TH), n);
    LLFE_KERNEL(ctx, "k_blur5");
    k_blur5<false><<<grid, 256, 0, ctx->stream>>>(src, dst, h, w, c);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

int launch_gray_blur5(llfe_ctx* ctx, const uint8_t* bgr, int n, int h, int w, uint8_t* dst) {
    if (n == 0 || h == 0 || w == 0) return LLFE_OK;
    // widths the streaming front kernel takes: its blurred-plane-only instance (registers instead of shared-memory tiles,
    // 128-bit loads, packed 16x2 arithmetic) is ~3x faster than the tile kernel below; "unfused" keeps the tile kernel
    // as the independent cross-check of the parity tests
    if (fused_supported(h, w) && !ctx->opt_unfused)
        return launch_fused(ctx, bgr, n, h, w, 0, 0, nullptr, nullptr, nullptr, nullptr, dst);
    dim3 grid(ceil_div(w, TWB), ceil_div(h, TH), n);
    LLFE_KERNEL(ctx, "k_gray_blur5");
    k_blur5<true><<<grid, 256, 0, ctx->stream>>>(bgr, dst, h, w, 1);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}
