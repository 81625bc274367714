// Palette front end (ColorExtractor.extract_colors, color_extractor.py:151, :224-225, :177):
//   BGR -> RGB, add int8 noise, clip, and np.unique(pixels, axis=0).
//
// np.unique(axis=0) sorts distinct RGB rows lexicographically == ascending
// key = R<<16 | G<<8 | B.  So: one 2^24-bit BITMAP per image (2 MiB, L2 resident),
// set with test-before-atomicOr, then an ordered compaction (popcount prefix).
// The popcount prefix doubles as a perfect hash (rank) of every present colour,
// which the optional counting pass uses to histogram pixels per unique colour.
#include "llfe_common.cuh"
#include "llfe_device.cuh"

namespace {

constexpr int BM_WORDS = 1 << 19;    // 2^24 bits
constexpr int BM_BLOCK = 1024;       // words per scan block
constexpr int BM_NBLK = BM_WORDS / BM_BLOCK;  // 512

__device__ __forceinline__ uint32_t noisy_key(uint32_t b, uint32_t g, uint32_t r, int nr, int ng, int nb) {
    int R = min(max((int)r + nr, 0), 255), G = min(max((int)g + ng, 0), 255), B = min(max((int)b + nb, 0), 255);
    return ((uint32_t)R << 16) | ((uint32_t)G << 8) | (uint32_t)B;
}

// MODE 0: set bitmap bits.  MODE 1: histogram pixels into hist[rank(key)].
// One thread per group of 8 consecutive pixels (the unit of the device noise generator).
template <int MODE>
__global__ void __launch_bounds__(256) k_color_pass(const uint8_t* __restrict__ bgr, size_t npix,
                                                    const int8_t* __restrict__ noise, uint64_t seed, int img0,
                                                    uint32_t* bitmap, const uint32_t* __restrict__ rank,
                                                    uint32_t* hist, int max_unique) {
    const int img = blockIdx.y;
    const uint8_t* s = bgr + (size_t)img * npix * 3;
    const int8_t* nz = noise ? noise + (size_t)img * npix * 3 : nullptr;
    uint32_t* bm = bitmap + (size_t)img * BM_WORDS;
    const uint32_t* rk = MODE == 1 ? rank + (size_t)img * BM_WORDS : nullptr;
    uint32_t* hs = MODE == 1 ? hist + (size_t)img * max_unique : nullptr;
    const size_t ngroups = (npix + 7) / 8;
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t g = blockIdx.x * (size_t)256 + threadIdx.x; g < ngroups; g += stride) {
        const size_t p0 = g * 8;
        const int np = (int)(npix - p0 < 8 ? npix - p0 : 8);
        uint8_t px[24];
        for (int k = 0; k < 3 * np; ++k) px[k] = s[3 * p0 + k];
        if (nz) {
            // injected noise is in RGB order, the image bytes in BGR order
            for (int j = 0; j < np; ++j) {
                const uint32_t key = noisy_key(px[3 * j], px[3 * j + 1], px[3 * j + 2], nz[3 * (p0 + j)],
                                               nz[3 * (p0 + j) + 1], nz[3 * (p0 + j) + 2]);
                px[3 * j] = (uint8_t)key;
                px[3 * j + 1] = (uint8_t)(key >> 8);
                px[3 * j + 2] = (uint8_t)(key >> 16);
            }
        } else {
            noise_apply_group(seed, ((uint64_t)(img0 + img) * npix + p0) >> 3, px, 3 * np);
        }
        for (int j = 0; j < np; ++j) {
            const uint32_t key = (uint32_t)px[3 * j] | ((uint32_t)px[3 * j + 1] << 8) | ((uint32_t)px[3 * j + 2] << 16);
            const uint32_t wi = key >> 5, bit = 1u << (key & 31);
            if (MODE == 0) {
                if (!(bm[wi] & bit)) atomicOr(&bm[wi], bit);  // stale reads only cost a redundant atomic
            } else {
                uint32_t idx = rk[wi] + __popc(bm[wi] & (bit - 1));
                if (idx < (uint32_t)max_unique) atomicAdd(&hs[idx], 1u);
            }
        }
    }
}

// per-1024-word block popcount
__global__ void __launch_bounds__(256) k_bm_blocksum(const uint32_t* __restrict__ bitmap, uint32_t* __restrict__ blocksum) {
    const int img = blockIdx.y, blk = blockIdx.x;
    const uint4* p = reinterpret_cast<const uint4*>(bitmap + (size_t)img * BM_WORDS + (size_t)blk * BM_BLOCK);
    uint4 v = p[threadIdx.x];
    uint32_t c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    c = warp_sum_u32(c);
    __shared__ uint32_t ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; ++i) t += ws[i];
        blocksum[(size_t)img * BM_NBLK + blk] = t;
    }
}

// exclusive scan of the 512 block sums of one image; also the image's unique count
__global__ void __launch_bounds__(BM_NBLK) k_bm_blockscan(uint32_t* __restrict__ blocksum, int32_t* __restrict__ count) {
    const int img = blockIdx.x, t = threadIdx.x;
    __shared__ uint32_t sh[BM_NBLK];
    uint32_t v = blocksum[(size_t)img * BM_NBLK + t];
    sh[t] = v;
    __syncthreads();
    for (int o = 1; o < BM_NBLK; o <<= 1) {
        uint32_t add = t >= o ? sh[t - o] : 0;
        __syncthreads();
        sh[t] += add;
        __syncthreads();
    }
    blocksum[(size_t)img * BM_NBLK + t] = sh[t] - v;
    if (t == BM_NBLK - 1) count[img] = (int32_t)sh[t];
}

// ordered emission of the set bits as keys; optionally the per-word rank table
__global__ void __launch_bounds__(256) k_bm_emit(const uint32_t* __restrict__ bitmap,
                                                 const uint32_t* __restrict__ blockofs, uint32_t* __restrict__ keys,
                                                 uint32_t* __restrict__ rank, int max_unique) {
    const int img = blockIdx.y, blk = blockIdx.x, t = threadIdx.x;
    const size_t wbase = (size_t)img * BM_WORDS + (size_t)blk * BM_BLOCK;
    uint4 v = reinterpret_cast<const uint4*>(bitmap + wbase)[t];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t c = __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
    // block exclusive scan of c over 256 threads
    __shared__ uint32_t wsum[8];
    uint32_t inc = c;
    const int lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t wofs = 0;
    for (int i = 0; i < warp; ++i) wofs += wsum[i];
    uint32_t pos = blockofs[(size_t)img * BM_NBLK + blk] + wofs + inc - c;
    uint32_t* out = keys + (size_t)img * max_unique;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t word = w[j];
        if (rank) rank[wbase + 4 * t + j] = pos;
        uint32_t base_key = (uint32_t)(((size_t)blk * BM_BLOCK + 4 * t + j) << 5);
        while (word) {
            int bpos = __ffs(word) - 1;
            word &= word - 1;
            if (pos < (uint32_t)max_unique) out[pos] = base_key | (uint32_t)bpos;
            ++pos;
        }
    }
}

}  // namespace

// workspace per image: bitmap (2 MiB) + rank (2 MiB, only with d_hist) + block sums
static size_t unique_ws_per_image(bool with_rank) {
    return (size_t)BM_WORDS * 4 * (with_rank ? 2 : 1) + WsCarver::need(BM_NBLK * 4);
}

int launch_unique_colors(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, const int8_t* d_noise, uint64_t seed,
                         int first_image, uint32_t* d_keys, uint32_t* d_hist, int32_t* d_count, int max_unique) {
    const size_t npix = (size_t)h * w;
    const bool with_rank = d_hist != nullptr;
    // chunk the batch so that the bitmaps of a chunk stay L2-sized (<= 64 MiB of bitmaps)
    const int chunk_max = with_rank ? 16 : 32;
    const int chunk = n < chunk_max ? n : chunk_max;
    void* ws;
    LLFE_TRY(llfe_workspace(ctx, unique_ws_per_image(with_rank) * chunk + 1024, &ws));
    WsCarver carve(ws);
    uint32_t* bitmap = carve.take<uint32_t>((size_t)BM_WORDS * chunk);
    uint32_t* rank = with_rank ? carve.take<uint32_t>((size_t)BM_WORDS * chunk) : nullptr;
    uint32_t* bsum = carve.take<uint32_t>((size_t)BM_NBLK * chunk);
    if (d_hist) LLFE_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)n * max_unique * sizeof(uint32_t), ctx->stream));
    size_t want = ceil_div_sz(npix, 256 * 8 * 2);
    unsigned gx = (unsigned)(want < 1 ? 1 : (want > 4096 ? 4096 : want));
    for (int i0 = 0; i0 < n; i0 += chunk) {
        const int m = (n - i0) < chunk ? (n - i0) : chunk;
        const uint8_t* src = d_bgr + (size_t)i0 * npix * 3;
        const int8_t* nz = d_noise ? d_noise + (size_t)i0 * npix * 3 : nullptr;
        LLFE_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)BM_WORDS * 4 * m, ctx->stream));
        LLFE_KERNEL(ctx, "k_color_bitmap");
        k_color_pass<0><<<dim3(gx, m), 256, 0, ctx->stream>>>(src, npix, nz, seed, first_image + i0, bitmap, nullptr, nullptr,
                                                              max_unique);
        LLFE_LAUNCHED(ctx);
        LLFE_TRY(launch_bitmap_compact(ctx, bitmap, bsum, m, d_keys + (size_t)i0 * max_unique, rank, d_count + i0,
                                       max_unique));
        if (d_hist) {
            LLFE_KERNEL(ctx, "k_color_count");
            k_color_pass<1><<<dim3(gx, m), 256, 0, ctx->stream>>>(src, npix, nz, seed, first_image + i0, bitmap, rank,
                                                                 d_hist + (size_t)i0 * max_unique, max_unique);
            LLFE_LAUNCHED(ctx);
        }
    }
    return LLFE_OK;
}

size_t bitmap_words_per_image() { return BM_WORDS; }
size_t bitmap_blocks_per_image() { return BM_NBLK; }

// bitmap (m images) -> sorted keys + counts; `rank` (optional) receives the per-word popcount prefix
int launch_bitmap_compact(llfe_ctx* ctx, const uint32_t* bitmap, uint32_t* bsum, int m, uint32_t* d_keys, uint32_t* rank,
                          int32_t* d_count, int max_unique) {
    LLFE_KERNEL(ctx, "k_bm_blocksum");
    k_bm_blocksum<<<dim3(BM_NBLK, m), 256, 0, ctx->stream>>>(bitmap, bsum);
    LLFE_LAUNCHED(ctx);
    LLFE_KERNEL(ctx, "k_bm_blockscan");
    k_bm_blockscan<<<m, BM_NBLK, 0, ctx->stream>>>(bsum, d_count);
    LLFE_LAUNCHED(ctx);
    LLFE_KERNEL(ctx, "k_bm_emit");
    k_bm_emit<<<dim3(BM_NBLK, m), 256, 0, ctx->stream>>>(bitmap, bsum, d_keys, rank, max_unique);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

extern "C" int llfe_unique_colors(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, const int8_t* d_noise,
                                  uint64_t seed, int first_image, uint32_t* d_keys, uint32_t* d_hist, int32_t* d_count,
                                  int max_unique) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_bgr != nullptr && d_keys != nullptr && d_count != nullptr);
    LLFE_CHECK_ARG(n >= 0 && n <= 65535 && h > 0 && w > 0 && max_unique > 0 && first_image >= 0);
    if (n == 0) return LLFE_OK;
    return launch_unique_colors(ctx, d_bgr, n, h, w, d_noise, seed, first_image, d_keys, d_hist, d_count, max_unique);
}
