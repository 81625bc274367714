// Per-pixel k-means building blocks for a row shard of ONE image (multi-GPU mode, BASELINE
// config 5): every rank runs `step` on its rows, the K x 4 uint64 accumulator is all-reduced
// (ncclSum over NVLink), then every rank runs the identical `update`.
//
// k_pixels_step streams the shard once per Lloyd iteration: 3 x 128-bit loads bring 16 BGR
// pixels per thread, bytes become exact floats with a byte-permute into 2^23 + x, and the
// nearest-centre search runs on PAIRS of pixels with packed f32x2 arithmetic (FFMA2: two
// IEEE-exact float32 operations per instruction; add.rn / mul.rn kept separate exactly like
// OpenCV's normL2Sqr, never fused).  Per-cluster sums are exact integers reduced with
// ballot + REDUX per warp, shared-memory tables per warp, one 64-bit atomic per table entry.
#include <string.h>

#include "llfe_common.cuh"
#include "k_kmeans_p2p.cuh"
#include "llfe_device.cuh"
#include "k_kmeans_shared.cuh"

namespace {

constexpr int PT = 256;
constexpr unsigned FULL = 0xffffffffu;

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
    return ((u64)__float_as_uint(hi) << 32) | (u64)__float_as_uint(lo);
}
__device__ __forceinline__ float lo_of(u64 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_of(u64 v) { return __uint_as_float((uint32_t)(v >> 32)); }
// two IEEE float32 operations per instruction (sm_100: FFMA2); results are the individually rounded
// add.rn / mul.rn of each half
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

__device__ __forceinline__ float fdist3(float r, float g, float b, const float* c) {
    float t0 = __fsub_rn(r, c[0]), t1 = __fsub_rn(g, c[1]), t2 = __fsub_rn(b, c[2]);
    float d = __fmul_rn(t0, t0);
    d = __fadd_rn(d, __fmul_rn(t1, t1));
    d = __fadd_rn(d, __fmul_rn(t2, t2));
    return d;
}

struct Acc {
    uint32_t v[PT / 32][KMAX][4];
};

// ballot + REDUX accumulation of one "row" of 32 pixels (one per lane) into the warp's table
__device__ __forceinline__ void accumulate_row(Acc& acc, int warp, int lane, int K, int bl, uint32_t r, uint32_t g,
                                               uint32_t b) {
    uint32_t todo = __ballot_sync(FULL, bl >= 0);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int k = __shfl_sync(FULL, bl, leader);
        const uint32_t m = __ballot_sync(FULL, bl == k);
        const bool in = bl == k;
        const uint32_t sr = __reduce_add_sync(FULL, in ? r : 0u);
        const uint32_t sg = __reduce_add_sync(FULL, in ? g : 0u);
        const uint32_t sb = __reduce_add_sync(FULL, in ? b : 0u);
        if (lane == 0) {
            acc.v[warp][k][0] += sr;
            acc.v[warp][k][1] += sg;
            acc.v[warp][k][2] += sb;
            acc.v[warp][k][3] += __popc(m);
        }
        todo &= ~m;
    }
}

__device__ __forceinline__ int nearest_scalar(float fr, float fg, float fb, const float (*s_c)[3], int K) {
    float bd = fdist3(fr, fg, fb, s_c[0]);
    int bl = 0;
    for (int k = 1; k < K; ++k) {
        float d = fdist3(fr, fg, fb, s_c[k]);
        if (d < bd) {
            bd = d;
            bl = k;
        }
    }
    return bl;
}

// nearest centre for every pixel + exact per-cluster sums
__global__ void __launch_bounds__(PT, 2) k_pixels_step(const uint8_t* __restrict__ bgr, size_t npix, int K,
                                                    const float* __restrict__ centers, unsigned long long* sums,
                                                    uint8_t* __restrict__ labels_out, int head,
                                                    const int32_t* __restrict__ state) {
    if (state && (state[1] | state[3])) return;   // converged, or waiting for the host (empty-cluster repair)
    __shared__ float s_c[KMAX][3];
    __shared__ u64 s_nc[KMAX][3];   // (-c, -c) pairs for the packed path
    __shared__ Acc s_acc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < K * 3; i += PT) {
        const float c = centers[i];
        (&s_c[0][0])[i] = c;
        (&s_nc[0][0])[i] = pack2(-c, -c);
    }
    for (int i = tid; i < (PT / 32) * KMAX * 4; i += PT) (&s_acc.v[0][0][0])[i] = 0u;
    __syncthreads();
    // ---- bulk: groups of 16 pixels = 48 bytes = 3 aligned 128-bit words, one group per thread ----
    const size_t nbulk = npix > (size_t)head ? (npix - head) / 16 : 0;
    const uint4* base = reinterpret_cast<const uint4*>(bgr + (size_t)head * 3);
    const size_t gstride = (size_t)gridDim.x * PT;
    for (size_t g0 = blockIdx.x * (size_t)PT; g0 < nbulk; g0 += gstride) {   // warp-uniform trip count
        const size_t g = g0 + tid;
        const bool ok = g < nbulk;
        uint32_t w[12];
        if (ok) {
            const uint4 a = ld_stream(base + 3 * g), b = ld_stream(base + 3 * g + 1), c = ld_stream(base + 3 * g + 2);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
            w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
            w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
        } else {
#pragma unroll
            for (int j = 0; j < 12; ++j) w[j] = 0u;
        }
        // unpack: pixel pair j = pixels (2j, 2j+1) = bytes 6j .. 6j+5 of the 48; every channel byte becomes
        // the exact float 2^23 + x - 2^23 (byte permute + one add), packed two pixels per 64-bit register
        u64 r2[8], g2[8], b2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float ch[6];
#pragma unroll
            for (int t = 0; t < 6; ++t) {
                const int bi = 6 * j + t;  // compile-time byte index in the 48
                const uint32_t m = __byte_perm(w[bi >> 2], 0x4B000000u, 0x7650u + (uint32_t)(bi & 3));
                ch[t] = __uint_as_float(m) - 8388608.0f;
            }
            // ch = {b0, g0, r0, b1, g1, r1}
            r2[j] = pack2(ch[2], ch[5]);
            g2[j] = pack2(ch[1], ch[4]);
            b2[j] = pack2(ch[0], ch[3]);
        }
        float bd[16];
        uint32_t lab16[4] = {0u, 0u, 0u, 0u};  // 16 labels, one byte each
        {
            const u64 n0 = s_nc[0][0], n1 = s_nc[0][1], n2 = s_nc[0][2];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const u64 t0 = add2(r2[j], n0), t1 = add2(g2[j], n1), t2 = add2(b2[j], n2);
                u64 d = mul2(t0, t0);
                d = add2(d, mul2(t1, t1));
                d = add2(d, mul2(t2, t2));
                bd[2 * j] = lo_of(d);
                bd[2 * j + 1] = hi_of(d);
            }
        }
        for (int k = 1; k < K; ++k) {
            const u64 n0 = s_nc[k][0], n1 = s_nc[k][1], n2 = s_nc[k][2];
            const uint32_t kk = (uint32_t)k;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const u64 t0 = add2(r2[j], n0), t1 = add2(g2[j], n1), t2 = add2(b2[j], n2);
                u64 d = mul2(t0, t0);
                d = add2(d, mul2(t1, t1));
                d = add2(d, mul2(t2, t2));
                const float d0 = lo_of(d), d1 = hi_of(d);
                // strict '<': the lowest index wins ties; label byte (2j) / (2j+1) of lab16
                if (d0 < bd[2 * j]) {
                    bd[2 * j] = d0;
                    lab16[j >> 1] = (lab16[j >> 1] & ~(0xffu << (16 * (j & 1)))) | (kk << (16 * (j & 1)));
                }
                if (d1 < bd[2 * j + 1]) {
                    bd[2 * j + 1] = d1;
                    lab16[j >> 1] = (lab16[j >> 1] & ~(0xff00u << (16 * (j & 1)))) | (kk << (16 * (j & 1) + 8));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int l0 = (int)((lab16[j >> 1] >> (16 * (j & 1))) & 0xffu), l1 = (int)((lab16[j >> 1] >> (16 * (j & 1) + 8)) & 0xffu);
            // the channel bytes again, as integers: byte 6j + t of the 48 (compile-time positions)
#define LLFE_BYTE(t) (__byte_perm(w[(6 * j + (t)) >> 2], 0u, 0x4440u + (uint32_t)((6 * j + (t)) & 3)))
            accumulate_row(s_acc, warp, lane, K, ok ? l0 : -1, LLFE_BYTE(2), LLFE_BYTE(1), LLFE_BYTE(0));
            accumulate_row(s_acc, warp, lane, K, ok ? l1 : -1, LLFE_BYTE(5), LLFE_BYTE(4), LLFE_BYTE(3));
#undef LLFE_BYTE
        }
        if (labels_out && ok) {
            uint8_t* lo = labels_out + head + 16 * g;
            if (((uintptr_t)lo & 15) == 0) {
                *reinterpret_cast<uint4*>(lo) = make_uint4(lab16[0], lab16[1], lab16[2], lab16[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) lo[j] = (uint8_t)(lab16[j >> 2] >> (8 * (j & 3)));
            }
        }
    }
    // ---- head (before the first aligned group) and tail pixels: scalar, block 0 only ----------------
    if (blockIdx.x == 0) {
        const size_t tail0 = (size_t)head + nbulk * 16;
        const size_t nrest = (npix < (size_t)head ? npix : (size_t)head) + (npix > tail0 ? npix - tail0 : 0);
        for (size_t q0 = 0; q0 < nrest; q0 += PT) {   // uniform trip count
            const size_t q = q0 + tid;
            int bl = -1;
            uint32_t b = 0, g = 0, r = 0;
            if (q < nrest) {
                const size_t hd = npix < (size_t)head ? npix : (size_t)head;
                const size_t p = q < hd ? q : tail0 + (q - hd);
                b = bgr[3 * p];
                g = bgr[3 * p + 1];
                r = bgr[3 * p + 2];
                bl = nearest_scalar((float)r, (float)g, (float)b, s_c, K);
                if (labels_out) labels_out[p] = (uint8_t)bl;
            }
            accumulate_row(s_acc, warp, lane, K, bl, r, g, b);
        }
    }
    __syncthreads();
    for (int i = tid; i < K * 4; i += PT) {
        unsigned long long t = 0;
        for (int w = 0; w < PT / 32; ++w) t += s_acc.v[w][i >> 2][i & 3];
        if (t) atomicAdd(&sums[i], t);
    }
}

// centres from (all-reduced) sums; shift; iteration bookkeeping (state: iter, done, n_empty)
//
// `copy_dst` (optional) receives the sums this update consumed and `zero_src` clears `sums` afterwards: with both,
// the per-rank accumulator can be all-reduced IN PLACE every iteration -- it is zero again before the next
// assignment adds to it, stays zero through the no-op iterations of a batch, and the totals survive in copy_dst.
__global__ void __launch_bounds__(32) k_pixels_update(int K, unsigned long long* sums, float* centers, int max_iter,
                                                      double eps2, int32_t* state, double* shift_out,
                                                      unsigned long long* copy_dst, int zero_src) {
    const int lane = threadIdx.x;   // one warp, lane = cluster (K <= 32)
    const int blocked = state[1] | state[3];
    const int it0 = state[0];
    __syncwarp();
    if (blocked) return;  // converged / frozen: later iterations of an unsynchronised batch are no-ops
    unsigned long long s[4] = {0ull, 0ull, 0ull, 1ull};
    if (lane < K) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s[j] = sums[4 * lane + j];
            if (copy_dst) {
                copy_dst[4 * lane + j] = s[j];
                if (zero_src) sums[4 * lane + j] = 0ull;
            }
        }
    }
    const int n_empty = __popc(__ballot_sync(FULL, s[3] == 0ull));
    if (n_empty) {
        if (lane == 0) {
            state[2] = n_empty;
            state[3] = 1;  // freeze: the host repairs the sums, clears state[2..3] and calls update again
        }
        return;
    }
    double sh = 0.0;
    if (lane < K) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float c = (float)((double)s[j] / (double)s[3]);
            const double t = (double)__fsub_rn(c, centers[3 * lane + j]);
            sh = __dadd_rn(sh, __dmul_rn(t, t));
            centers[3 * lane + j] = c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sh = fmax(sh, __shfl_xor_sync(FULL, sh, o));
    if (lane == 0) {
        const int it = it0 + 1;
        const int last_it = max_iter > 2 ? max_iter : 2;
        state[0] = it;
        state[2] = 0;
        state[1] = (it == last_it) || (it0 > 0 && sh <= eps2);
        if (shift_out) *shift_out = sh;
    }
}

// ---- fused all-reduce + centre update over peer memory (config 5, N > 1): see k_kmeans_p2p.cuh ------------------------
__global__ void __launch_bounds__(128) k_pixels_update_p2p(int K, unsigned long long* sums, P2PMailbox* const* peers, int rank,
                                                           int world, float* centers, int max_iter, double eps2,
                                                           int32_t* state, double* shift_out, unsigned long long* totals) {
    __shared__ unsigned long long s_tot[KMAX * 4];
    p2p_exchange_and_update(K, sums, peers, rank, world, centers, max_iter, eps2, state, shift_out, totals, s_tot);
}

// zero the per-rank accumulator for the next iteration -- unless the loop is converged or frozen
__global__ void k_pixels_zero(int K, unsigned long long* sums, const int32_t* __restrict__ state) {
    if (state && (state[1] | state[3])) return;
    if (threadIdx.x < K * 4) sums[threadIdx.x] = 0ull;
}

}  // namespace

extern "C" int llfe_kmeans_pixels_zero(llfe_ctx* ctx, int k, uint64_t* d_sums_counts, const int32_t* d_state_or_null) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_sums_counts != nullptr && k >= 1 && k <= KMAX);
    LLFE_KERNEL(ctx, "k_pixels_zero");
    k_pixels_zero<<<1, 128, 0, ctx->stream>>>(k, (unsigned long long*)d_sums_counts, d_state_or_null);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

extern "C" int llfe_kmeans_pixels_step(llfe_ctx* ctx, const uint8_t* d_bgr, size_t n_pixels, int k,
                                       const float* d_centers, uint64_t* d_sums_counts, uint8_t* d_labels_or_null,
                                       const int32_t* d_state_or_null) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_bgr != nullptr && d_centers != nullptr && d_sums_counts != nullptr);
    LLFE_CHECK_ARG(k >= 1 && k <= KMAX);
    if (n_pixels == 0) return LLFE_OK;
    // pixels before the first 16-byte boundary that is also a pixel boundary: 3 h = -addr (mod 16)
    const int head = (int)(((16 - ((uintptr_t)d_bgr & 15)) & 15) * 11 % 16);
    const size_t groups = n_pixels / 16 + 1;
    size_t want = ceil_div_sz(groups, PT);
    size_t cap = (size_t)ctx->sm_count * 8;
    unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
    LLFE_KERNEL(ctx, "k_pixels_step");
    k_pixels_step<<<grid, PT, 0, ctx->stream>>>(d_bgr, n_pixels, k, d_centers, (unsigned long long*)d_sums_counts,
                                                d_labels_or_null, head, d_state_or_null);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

extern "C" int llfe_kmeans_update(llfe_ctx* ctx, int k, uint64_t* d_sums_counts, float* d_centers, int max_iter,
                                  double eps, int32_t* d_state, double* d_shift, uint64_t* d_consumed_or_null,
                                  int zero_sums) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_sums_counts != nullptr && d_centers != nullptr && d_state != nullptr);
    LLFE_CHECK_ARG(k >= 1 && k <= KMAX && max_iter >= 1 && (!zero_sums || d_consumed_or_null != nullptr));
    LLFE_CHECK_ARG(d_consumed_or_null != d_sums_counts);
    LLFE_KERNEL(ctx, "k_pixels_update");
    k_pixels_update<<<1, 32, 0, ctx->stream>>>(k, (unsigned long long*)d_sums_counts, d_centers, max_iter, eps * eps,
                                               d_state, d_shift, (unsigned long long*)d_consumed_or_null, zero_sums);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

extern "C" size_t llfe_p2p_mailbox_bytes(void) { return sizeof(P2PMailbox); }

extern "C" int llfe_ipc_export(llfe_ctx* ctx, void* d_ptr, uint8_t* handle64) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_ptr != nullptr && handle64 != nullptr);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    cudaIpcMemHandle_t h;
    LLFE_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle64, &h, 64);
    return LLFE_OK;
}

extern "C" int llfe_ipc_open(llfe_ctx* ctx, const uint8_t* handle64, void** d_peer_out) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(handle64 != nullptr && d_peer_out != nullptr);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    LLFE_CUDA(cudaIpcOpenMemHandle(d_peer_out, h, cudaIpcMemLazyEnablePeerAccess));
    return LLFE_OK;
}

extern "C" int llfe_ipc_close(llfe_ctx* ctx, void* d_peer) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_peer != nullptr);
    LLFE_CUDA(cudaIpcCloseMemHandle(d_peer));
    return LLFE_OK;
}

extern "C" int llfe_kmeans_update_p2p(llfe_ctx* ctx, int k, uint64_t* d_partial_sums, void* const* d_mailboxes, int rank,
                                      int world, float* d_centers, int max_iter, double eps, int32_t* d_state,
                                      double* d_shift, uint64_t* d_totals) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_partial_sums != nullptr && d_mailboxes != nullptr && d_centers != nullptr && d_state != nullptr &&
                   d_totals != nullptr && d_totals != d_partial_sums);
    LLFE_CHECK_ARG(k >= 1 && k <= KMAX && max_iter >= 1 && world >= 1 && world <= P2P_MAXW && rank >= 0 && rank < world);
    LLFE_KERNEL(ctx, "k_pixels_update_p2p");
    k_pixels_update_p2p<<<1, 128, 0, ctx->stream>>>(k, (unsigned long long*)d_partial_sums, (P2PMailbox* const*)d_mailboxes,
                                                    rank, world, d_centers, max_iter, eps * eps, d_state, d_shift,
                                                    (unsigned long long*)d_totals);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}
