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

// the 8 pixels of 6 packed words as 24-bit values (byte 0 first): B | G<<8 | R<<16 for image words
__device__ __forceinline__ void unpack24(const uint32_t* w, uint32_t* o) {
    o[0] = w[0] & 0xffffffu;
    o[1] = __byte_perm(w[0], w[1], 0x0543) & 0xffffffu;
    o[2] = __byte_perm(w[1], w[2], 0x0432) & 0xffffffu;
    o[3] = w[2] >> 8;
    o[4] = w[3] & 0xffffffu;
    o[5] = __byte_perm(w[3], w[4], 0x0543) & 0xffffffu;
    o[6] = __byte_perm(w[4], w[5], 0x0432) & 0xffffffu;
    o[7] = w[5] >> 8;
}

// 24 bytes at p (pixel-aligned, any byte alignment) as 6 little-endian words; bytes beyond `nbytes` read as zero
__device__ __forceinline__ void load24(const uint8_t* __restrict__ p, int nbytes, bool aligned8, uint32_t* w) {
    if (aligned8 && nbytes == 24) {
        const uint2* q = reinterpret_cast<const uint2*>(p);
        const uint2 a = q[0], b = q[1], c = q[2];
        w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x; w[5] = c.y;
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (4 * k + j < nbytes) v |= (uint32_t)p[4 * k + j] << (8 * j);
            w[k] = v;
        }
    }
}

constexpr int CP_WARPS = 8;

// The colour pass: BGR -> RGB key, noise, clip, then MODE 0: set the key's bit in the image's bitmap (test before
// atomicOr: a stale read only costs a redundant atomic); MODE 1: histogram the pixel into hist[rank(key)].
// A warp owns a block of 256 consecutive pixels (the unit of the device noise, llfe_device.cuh), lane L its pixels
// 8L .. 8L+7: 768 contiguous bytes per warp and step, 24 per lane.  Pointwise: no halo, no barrier.
template <int MODE>
__global__ void __launch_bounds__(CP_WARPS * 32) k_color_pass(const uint8_t* __restrict__ bgr, size_t npix,
                                                              const int8_t* __restrict__ noise, uint64_t seed, int img0,
                                                              uint32_t* bitmap, const uint32_t* __restrict__ rank,
                                                              uint32_t* hist, int max_unique) {
    __shared__ uint2 s_bytes[CP_WARPS][96];   // 768 bytes per warp
    const int img = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint8_t* s = bgr + (size_t)img * npix * 3;
    const int8_t* nz = noise ? noise + (size_t)img * npix * 3 : nullptr;
    uint32_t* bm = bitmap + (size_t)img * BM_WORDS;
    const uint32_t* rk = MODE == 1 ? rank + (size_t)img * BM_WORDS : nullptr;
    uint32_t* hs = MODE == 1 ? hist + (size_t)img * max_unique : nullptr;
    const bool aligned8 = (((uintptr_t)s) & 7) == 0 && (!nz || (((uintptr_t)nz) & 7) == 0);
    const bool aligned16 = (((uintptr_t)s) & 15) == 0;
    const size_t nblocks = (npix + LLFE_NOISE_BLOCK_PX - 1) / LLFE_NOISE_BLOCK_PX;
    const size_t nfull = npix / LLFE_NOISE_BLOCK_PX;     // blocks with all 256 pixels
    uint2* mine = &s_bytes[warp][3 * lane];
    uint4* blk4 = reinterpret_cast<uint4*>(&s_bytes[warp][0]);   // the block as 48 x 16 bytes
    const uint4* src4 = reinterpret_cast<const uint4*>(s);
    const size_t stride = (size_t)gridDim.x * CP_WARPS;
    size_t b = (size_t)blockIdx.x * CP_WARPS + warp;
    // Device-noise path on full, 16-byte aligned blocks: the block's 768 bytes arrive as 1.5 coalesced 128-bit loads per
    // lane and go straight into the warp's shared-memory copy (where the noise is applied); the loads of the warp's NEXT
    // block are issued before this block is processed.
    const bool stream_ok = !nz && aligned16;
    uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
    if (stream_ok && b < nfull) {
        v0 = ld_stream(src4 + 48 * b + lane);
        if (lane < 16) v1 = ld_stream(src4 + 48 * b + 32 + lane);
    }
    for (; b < nblocks; b += stride) {
        const size_t p0 = b * LLFE_NOISE_BLOCK_PX + 8 * lane;
        const int np = p0 >= npix ? 0 : (int)(npix - p0 < 8 ? npix - p0 : 8);
        uint32_t key[8];
        if (nz) {
            // injected noise (the reference's tensor) is in RGB order, the image bytes in BGR order
            uint32_t w[6], nw[6], n24[8];
            load24(s + 3 * p0, 3 * np, aligned8, w);
            load24(reinterpret_cast<const uint8_t*>(nz) + 3 * p0, 3 * np, aligned8, nw);
            unpack24(w, key);
            unpack24(nw, n24);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (n24[j]) {
                    const uint32_t k = key[j];
                    key[j] = noisy_key(k & 255u, (k >> 8) & 255u, k >> 16, (int)(int8_t)(n24[j] & 255u),
                                       (int)(int8_t)((n24[j] >> 8) & 255u), (int)(int8_t)(n24[j] >> 16));
                }
        } else {
            if (stream_ok && b < nfull) {
                blk4[lane] = v0;
                if (lane < 16) blk4[32 + lane] = v1;
                const size_t bn = b + stride;
                if (bn < nfull) {
                    v0 = ld_stream(src4 + 48 * bn + lane);
                    if (lane < 16) v1 = ld_stream(src4 + 48 * bn + 32 + lane);
                }
            } else {
                uint32_t w[6];
                load24(s + 3 * p0, 3 * np, aligned8, w);
                mine[0] = make_uint2(w[0], w[1]);
                mine[1] = make_uint2(w[2], w[3]);
                mine[2] = make_uint2(w[4], w[5]);
            }
            __syncwarp();
            const size_t left = npix - b * LLFE_NOISE_BLOCK_PX;
            noise_apply_block(noise_block_base(seed, (uint32_t)(img0 + img), (uint32_t)b),
                              reinterpret_cast<uint8_t*>(&s_bytes[warp][0]),
                              3 * (int)(left < LLFE_NOISE_BLOCK_PX ? left : LLFE_NOISE_BLOCK_PX), lane);
            __syncwarp();
            const uint2 a = mine[0], c = mine[1], d = mine[2];
            const uint32_t v[6] = {a.x, a.y, c.x, c.y, d.x, d.y};
            unpack24(v, key);
            __syncwarp();   // the next step overwrites the block's bytes
        }
        if (MODE == 0) {
            uint32_t val[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) val[j] = j < np ? bm[key[j] >> 5] : 0xffffffffu;   // all loads in flight first
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (!(__funnelshift_r(val[j], 0u, key[j]) & 1u)) atomicOr(&bm[key[j] >> 5], 1u << (key[j] & 31u));
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j >= np) break;
                const uint32_t wi = key[j] >> 5, bit = 1u << (key[j] & 31u);
                const uint32_t idx = rk[wi] + __popc(bm[wi] & (bit - 1));
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

// CTAs per image: about two full waves of the machine over the m images of the launch (6 CTAs of 8 warps per SM), so
// that a warp walks a dozen blocks with its next block's loads in flight, but never more CTAs than blocks
static unsigned color_pass_grid(llfe_ctx* ctx, size_t npix, int m) {
    const size_t blocks = ceil_div_sz(npix, (size_t)LLFE_NOISE_BLOCK_PX * CP_WARPS);
    const size_t slots = (size_t)(ctx->sm_count > 0 ? ctx->sm_count : 148) * 6 * 2;
    size_t want = ceil_div_sz(slots, (size_t)(m > 0 ? m : 1));
    if (want > blocks) want = blocks;
    return (unsigned)(want < 1 ? 1 : want);
}

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
    const unsigned gx = color_pass_grid(ctx, npix, chunk);
    for (int i0 = 0; i0 < n; i0 += chunk) {
        const int m = (n - i0) < chunk ? (n - i0) : chunk;
        const uint8_t* src = d_bgr + (size_t)i0 * npix * 3;
        const int8_t* nz = d_noise ? d_noise + (size_t)i0 * npix * 3 : nullptr;
        LLFE_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)BM_WORDS * 4 * m, ctx->stream));
        LLFE_KERNEL(ctx, "k_color_pass");
        k_color_pass<0><<<dim3(gx, m), CP_WARPS * 32, 0, ctx->stream>>>(src, npix, nz, seed, first_image + i0, bitmap, nullptr, nullptr,
                                                              max_unique);
        LLFE_LAUNCHED(ctx);
        LLFE_TRY(launch_bitmap_compact(ctx, bitmap, bsum, m, d_keys + (size_t)i0 * max_unique, rank, d_count + i0,
                                       max_unique));
        if (d_hist) {
            LLFE_KERNEL(ctx, "k_color_count");
            k_color_pass<1><<<dim3(gx, m), CP_WARPS * 32, 0, ctx->stream>>>(src, npix, nz, seed, first_image + i0, bitmap, rank,
                                                                 d_hist + (size_t)i0 * max_unique, max_unique);
            LLFE_LAUNCHED(ctx);
        }
    }
    return LLFE_OK;
}

// set the colour bits of m images in `bitmap` (already zeroed); image i of this call is image first_image + i
// of the caller's numbering for the device noise
int launch_color_bitmap(llfe_ctx* ctx, const uint8_t* d_bgr, int m, int h, int w, const int8_t* d_noise, uint64_t seed,
                        int first_image, uint32_t* bitmap) {
    const size_t npix = (size_t)h * w;
    const unsigned gx = color_pass_grid(ctx, npix, m);
    LLFE_KERNEL(ctx, "k_color_pass");
    k_color_pass<0><<<dim3(gx, m), CP_WARPS * 32, 0, ctx->stream>>>(d_bgr, npix, d_noise, seed, first_image, bitmap, nullptr,
                                                                    nullptr, 0);
    LLFE_LAUNCHED(ctx);
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
