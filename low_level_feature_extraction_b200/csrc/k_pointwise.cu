// Pointwise ops: BGR->gray (Q15), BGR->RGB, convertScaleAbs (as a 256-entry LUT).
#include "llfe_common.cuh"
#include "llfe_device.cuh"

// ---------------------------------------------------------------------------
// BGR2GRAY.  16 pixels (48 bytes in, 16 bytes out) per thread on the aligned
// path: three 128-bit loads, eight IDP.2A per 4 pixels, one 128-bit store.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bgr2gray_vec(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                      size_t ngroups) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= ngroups) return;
    uint4 a = ld_stream(src + 3 * i), b = ld_stream(src + 3 * i + 1), c = ld_stream(src + 3 * i + 2);
    uint4 o;
    o.x = gray4_packed(a.x, a.y, a.z);
    o.y = gray4_packed(a.w, b.x, b.y);
    o.z = gray4_packed(b.z, b.w, c.x);
    o.w = gray4_packed(c.y, c.z, c.w);
    dst[i] = o;
}

__global__ void k_bgr2gray_scalar(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t first,
                                  size_t npix) {
    size_t i = first + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= npix) return;
    dst[i] = gray_px(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
}

int launch_bgr2gray(llfe_ctx* ctx, const uint8_t* bgr, size_t npix, uint8_t* gray) {
    if (npix == 0) return LLFE_OK;
    size_t nvec = 0;
    if (((uintptr_t)bgr % 16 == 0) && ((uintptr_t)gray % 16 == 0)) nvec = npix / 16;
    if (nvec) {
        LLFE_KERNEL(ctx, "k_bgr2gray_vec");
        k_bgr2gray_vec<<<(unsigned)ceil_div_sz(nvec, 256), 256, 0, ctx->stream>>>((const uint4*)bgr, (uint4*)gray, nvec);
        LLFE_LAUNCHED(ctx);
    }
    size_t done = nvec * 16;
    if (done < npix) {
        size_t rest = npix - done;
        LLFE_KERNEL(ctx, "k_bgr2gray_scalar");
        k_bgr2gray_scalar<<<(unsigned)ceil_div_sz(rest, 256), 256, 0, ctx->stream>>>(bgr, gray, done, npix);
        LLFE_LAUNCHED(ctx);
    }
    return LLFE_OK;
}

// ---------------------------------------------------------------------------
// BGR2RGB: byte swap inside each 3-byte pixel (4 pixels = 3 words per step).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bgr2rgb_vec(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                     size_t ngroups) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= ngroups) return;
    uint32_t w0 = src[3 * i], w1 = src[3 * i + 1], w2 = src[3 * i + 2];
    // bytes: b0 g0 r0 b1 | g1 r1 b2 g2 | r2 b3 g3 r3  ->  r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
    dst[3 * i] = __byte_perm(w0, w1, 0x5012);
    dst[3 * i + 1] = __byte_perm(__byte_perm(w0, w1, 0x7034), w2, 0x3410);  // g1 b1 r2 g2
    dst[3 * i + 2] = __byte_perm(w1, w2, 0x5672);
}

__global__ void k_bgr2rgb_scalar(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t first, size_t npix) {
    size_t i = first + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= npix) return;
    uint8_t b = src[3 * i], g = src[3 * i + 1], r = src[3 * i + 2];
    dst[3 * i] = r;
    dst[3 * i + 1] = g;
    dst[3 * i + 2] = b;
}

int launch_bgr2rgb(llfe_ctx* ctx, const uint8_t* bgr, size_t npix, uint8_t* rgb) {
    if (npix == 0) return LLFE_OK;
    size_t nvec = 0;
    if (((uintptr_t)bgr % 4 == 0) && ((uintptr_t)rgb % 4 == 0)) nvec = npix / 4;
    if (nvec) {
        LLFE_KERNEL(ctx, "k_bgr2rgb_vec");
        k_bgr2rgb_vec<<<(unsigned)ceil_div_sz(nvec, 256), 256, 0, ctx->stream>>>((const uint32_t*)bgr, (uint32_t*)rgb,
                                                                                nvec);
        LLFE_LAUNCHED(ctx);
    }
    size_t done = nvec * 4;
    if (done < npix) {
        LLFE_KERNEL(ctx, "k_bgr2rgb_scalar");
        k_bgr2rgb_scalar<<<(unsigned)ceil_div_sz(npix - done, 256), 256, 0, ctx->stream>>>(bgr, rgb, done, npix);
        LLFE_LAUNCHED(ctx);
    }
    return LLFE_OK;
}

// ---------------------------------------------------------------------------
// convertScaleAbs(alpha=a1, beta=0) [then (alpha=a2, beta=0)] as one LUT pass.
// lut[v] = sat_u8(rint(|f32(v) * f32(a)|)), round-half-even (cvRound).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t csa(uint32_t v, float a) {
    float f = fabsf(__fmul_rn((float)v, a));
    int r = __float2int_rn(fminf(f, 1e6f));
    return (uint32_t)min(max(r, 0), 255);
}

__global__ void __launch_bounds__(256) k_lut2(const uint8_t* __restrict__ src, size_t count, float a1, float a2,
                                              int single, uint8_t* __restrict__ dst) {
    __shared__ uint8_t lut[256];
    {
        uint32_t v = csa(threadIdx.x, a1);
        if (!single) v = csa(v, a2);
        lut[threadIdx.x] = (uint8_t)v;
    }
    __syncthreads();
    size_t nvec = (((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0)) ? count / 16 : 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (size_t i = tid; i < nvec; i += stride) {
        uint4 v = ld_stream(reinterpret_cast<const uint4*>(src) + i);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t x = w[j];
            w[j] = lut[x & 255] | (lut[(x >> 8) & 255] << 8) | (lut[(x >> 16) & 255] << 16) | (lut[x >> 24] << 24);
        }
        reinterpret_cast<uint4*>(dst)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (size_t i = nvec * 16 + tid; i < count; i += stride) dst[i] = lut[src[i]];
}

int launch_lut2(llfe_ctx* ctx, const uint8_t* src, size_t count, float a1, float a2, int single, uint8_t* dst) {
    if (count == 0) return LLFE_OK;
    size_t want = ceil_div_sz(count, 16 * 256);
    unsigned grid = (unsigned)(want < (size_t)ctx->sm_count * 16 ? (want ? want : 1) : (size_t)ctx->sm_count * 16);
    LLFE_KERNEL(ctx, "k_lut2");
    k_lut2<<<grid, 256, 0, ctx->stream>>>(src, count, a1, a2, single, dst);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}
