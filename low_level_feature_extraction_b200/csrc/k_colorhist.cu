// Colour-histogram form of the per-pixel k-means of ONE oversized image (BASELINE config 5).
//
// A u8 image has at most 2^24 distinct colours, and Lloyd's update only needs exact INTEGER sums
// (SURVEY.md 8(e)): sum over pixels of x = sum over distinct colours of count * x.  So the rank's
// rows are streamed ONCE into a 2^24-bin count table keyed (R << 16) + (G << 8) + B -- the key
// order is np.unique's lexicographic (R, G, B) order -- the table is all-reduced once, compacted
// into (key, count) entries, and every Lloyd iteration then touches the distinct colours only:
// the same labels, the same exact sums, the same centres as the per-pixel pass of k_pixels.cu,
// without re-reading the image.  Per-pixel labels, when asked for, are one lookup pass through a
// 16 MiB colour -> label table.
//
// Work split across ranks: the table is cut into blocks of HB keys and block b belongs to part
// b % parts, so each rank iterates over an interleaved share of the colour space and the K x 4
// accumulator is all-reduced per iteration exactly as in the per-pixel form.
#include "llfe_common.cuh"
#include "llfe_device.cuh"
#include "k_kmeans_shared.cuh"
#include "k_kmeans_p2p.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int HT = 256;
constexpr unsigned FULL = 0xffffffffu;
constexpr int HB = 2048;                  // keys per compaction block
constexpr int NBLK = (1 << 24) / HB;      // 8192 blocks
constexpr int EPT = HB / HT;              // table entries per thread in the compaction kernels

typedef unsigned long long u64;

// 16 pixels = 48 bytes in 12 words -> 16 keys; a BGR pixel's three bytes ARE the little-endian key
__device__ __forceinline__ void keys16(const uint32_t (&w)[13], uint32_t (&key)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int bi = 3 * j;
        key[j] = __funnelshift_r(w[bi >> 2], w[(bi >> 2) + 1], 8 * (bi & 3)) & 0xffffffu;
    }
}

__device__ __forceinline__ void load48(const uint4* p, uint32_t (&w)[13]) {
    const uint4 a = ld_stream(p), b = ld_stream(p + 1), c = ld_stream(p + 2);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
    w[12] = 0u;
}

// count table: hist[key] += 1 for every pixel (runs of equal keys inside a thread's 16 pixels are merged)
__global__ void __launch_bounds__(HT) k_hist_count(const uint8_t* __restrict__ bgr, size_t npix, int head,
                                                   uint32_t* __restrict__ hist) {
    const size_t nbulk = npix > (size_t)head ? (npix - head) / 16 : 0;
    const uint4* base = reinterpret_cast<const uint4*>(bgr + (size_t)head * 3);
    const size_t gstride = (size_t)gridDim.x * HT;
    for (size_t g = blockIdx.x * (size_t)HT + threadIdx.x; g < nbulk; g += gstride) {
        uint32_t w[13], key[16];
        load48(base + 3 * g, w);
        keys16(w, key);
        uint32_t cur = key[0], run = 1;
#pragma unroll
        for (int j = 1; j < 16; ++j) {
            if (key[j] == cur) {
                ++run;
            } else {
                atomicAdd(hist + cur, run);
                cur = key[j];
                run = 1;
            }
        }
        atomicAdd(hist + cur, run);
    }
    if (blockIdx.x == 0) {   // pixels before the first aligned group and after the last one
        const size_t hd = npix < (size_t)head ? npix : (size_t)head;
        const size_t tail0 = (size_t)head + nbulk * 16;
        const size_t nrest = hd + (npix > tail0 ? npix - tail0 : 0);
        for (size_t q = threadIdx.x; q < nrest; q += HT) {
            const size_t p = q < hd ? q : tail0 + (q - hd);
            atomicAdd(hist + (((uint32_t)bgr[3 * p + 2] << 16) | ((uint32_t)bgr[3 * p + 1] << 8) | bgr[3 * p]), 1u);
        }
    }
}

// ---- ordered compaction of the non-empty bins of the blocks that belong to `part` --------------
// packed != 0: `hist` is this part's share only, its blocks back to back (block b of the table at slot b / parts)
__global__ void __launch_bounds__(HT) k_hist_blockcount(const uint32_t* __restrict__ hist, int part, int parts, int packed,
                                                        uint32_t* __restrict__ bcount) {
    const int b = blockIdx.x;
    int c = 0;
    if (b % parts == part) {
        const uint4* p = reinterpret_cast<const uint4*>(hist + (size_t)(packed ? b / parts : b) * HB) + threadIdx.x * (EPT / 4);
#pragma unroll
        for (int i = 0; i < EPT / 4; ++i) {
            const uint4 v = p[i];
            c += (v.x != 0u) + (v.y != 0u) + (v.z != 0u) + (v.w != 0u);
        }
    }
    __shared__ int s[HT / 32];
    const int ws = __reduce_add_sync(FULL, c);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = ws;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < HT / 32; ++i) t += s[i];
        bcount[b] = (uint32_t)t;
    }
}

// exclusive scan of the NBLK block counts (one CTA), total -> *n_out
__global__ void __launch_bounds__(1024) k_hist_blockscan(const uint32_t* __restrict__ bcount,
                                                         uint32_t* __restrict__ boffs, int32_t* n_out) {
    constexpr int PER = NBLK / 1024;
    __shared__ uint32_t s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t v[PER], t = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        v[i] = bcount[tid * PER + i];
        t += v[i];
    }
    uint32_t inc = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t x = s_w[lane], y = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(FULL, y, o);
            if (lane >= o) y += n;
        }
        s_w[lane] = y - x;   // exclusive warp offsets
        if (lane == 31) *n_out = (int32_t)y;
    }
    __syncthreads();
    uint32_t run = s_w[warp] + inc - t;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        boffs[tid * PER + i] = run;
        run += v[i];
    }
}

__global__ void __launch_bounds__(HT) k_hist_emit(const uint32_t* __restrict__ hist, int part, int parts, int packed,
                                                  const uint32_t* __restrict__ boffs, uint32_t* __restrict__ keys,
                                                  uint32_t* __restrict__ counts, size_t cap) {
    const int b = blockIdx.x;
    if (b % parts != part) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t v[EPT];
    const uint4* p = reinterpret_cast<const uint4*>(hist + (size_t)(packed ? b / parts : b) * HB) + tid * (EPT / 4);
    int c = 0;
#pragma unroll
    for (int i = 0; i < EPT / 4; ++i) {
        const uint4 q = p[i];
        v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
        c += (q.x != 0u) + (q.y != 0u) + (q.z != 0u) + (q.w != 0u);
    }
    __shared__ int s_w[HT / 32];
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int woff = 0;
    for (int i = 0; i < warp; ++i) woff += s_w[i];
    size_t pos = (size_t)boffs[b] + woff + inc - c;
    const uint32_t key0 = (uint32_t)b * HB + tid * EPT;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        if (v[i] != 0u) {
            if (pos < cap) {
                keys[pos] = key0 + i;
                counts[pos] = v[i];
            }
            ++pos;
        }
    }
}

__device__ __forceinline__ float fdist3(float r, float g, float b, const float* c) {
    const float t0 = __fsub_rn(r, c[0]), t1 = __fsub_rn(g, c[1]), t2 = __fsub_rn(b, c[2]);
    float d = __fmul_rn(t0, t0);
    d = __fadd_rn(d, __fmul_rn(t1, t1));
    d = __fadd_rn(d, __fmul_rn(t2, t2));
    return d;
}

// two IEEE float32 operations per instruction (sm_100 FADD2 / FMUL2); each half is rounded on its own
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    return ((u64)__float_as_uint(hi) << 32) | (u64)__float_as_uint(lo);
}
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

constexpr int EPS_T = 4;   // entries per thread per trip of k_hist_step (two f32x2 pairs)

// one Lloyd assignment over (key, count) entries: cv2's float32 distance and first-minimum rule per
// distinct colour (on PAIRS of colours with packed f32x2 arithmetic, x + (-c) == x - c exactly), exact u64
// sums weighted by the pixel counts
// The body of one assignment, called by every thread of a block of HT threads; centres are read through L2 (the
// persistent kernel below rewrites them between its iterations).
__device__ __forceinline__ void hist_step_body(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ counts, size_t n,
                                               int K, const float* centers, u64* sums, uint8_t* __restrict__ labels_out,
                                               u64 (*s_nc)[3], u64 (*s_acc)[KMAX][4]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < K * 3; i += HT) {
        const float c = __ldcg(&centers[i]);
        (&s_nc[0][0])[i] = pack2(-c, -c);
    }
    for (int i = tid; i < (HT / 32) * KMAX * 4; i += HT) (&s_acc[0][0][0])[i] = 0ull;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * HT * EPS_T;
    for (size_t i0 = blockIdx.x * (size_t)(HT * EPS_T); i0 < n; i0 += stride) {   // block-uniform trip count
        uint32_t key[EPS_T], cnt[EPS_T];
        bool ok[EPS_T];
#pragma unroll
        for (int j = 0; j < EPS_T; ++j) {
            const size_t i = i0 + (size_t)j * HT + tid;
            ok[j] = i < n;
            key[j] = ok[j] ? keys[i] : 0u;
            cnt[j] = ok[j] ? counts[i] : 0u;
        }
        u64 r2[EPS_T / 2], g2[EPS_T / 2], b2[EPS_T / 2];
#pragma unroll
        for (int p = 0; p < EPS_T / 2; ++p) {
            r2[p] = pack2((float)(key[2 * p] >> 16), (float)(key[2 * p + 1] >> 16));
            g2[p] = pack2((float)((key[2 * p] >> 8) & 0xffu), (float)((key[2 * p + 1] >> 8) & 0xffu));
            b2[p] = pack2((float)(key[2 * p] & 0xffu), (float)(key[2 * p + 1] & 0xffu));
        }
        float bd[EPS_T];
        int bl[EPS_T];
        for (int k = 0; k < K; ++k) {
            const u64 n0 = s_nc[k][0], n1 = s_nc[k][1], n2 = s_nc[k][2];
#pragma unroll
            for (int p = 0; p < EPS_T / 2; ++p) {
                const u64 t0 = add2(r2[p], n0), t1 = add2(g2[p], n1), t2 = add2(b2[p], n2);
                u64 d = mul2(t0, t0);
                d = add2(d, mul2(t1, t1));
                d = add2(d, mul2(t2, t2));
                const float d0 = __uint_as_float((uint32_t)d), d1 = __uint_as_float((uint32_t)(d >> 32));
                if (k == 0 || d0 < bd[2 * p]) {   // strict: the lowest index wins ties
                    bd[2 * p] = d0;
                    bl[2 * p] = k;
                }
                if (k == 0 || d1 < bd[2 * p + 1]) {
                    bd[2 * p + 1] = d1;
                    bl[2 * p + 1] = k;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < EPS_T; ++j) {
            if (labels_out && ok[j]) labels_out[i0 + (size_t)j * HT + tid] = (uint8_t)bl[j];
            const int lab = ok[j] ? bl[j] : -1;
            const uint32_t r = key[j] >> 16, g = (key[j] >> 8) & 0xffu, b = key[j] & 0xffu;
            // counts reach 2^28: reduce the low and high 16 bits of count * channel separately (REDUX is 32-bit)
            const uint32_t cl = cnt[j] & 0xffffu, ch = cnt[j] >> 16;
            uint32_t todo = __ballot_sync(FULL, lab >= 0);
            while (todo) {
                const int k = __shfl_sync(FULL, lab, __ffs(todo) - 1);
                const bool in = lab == k;
                const uint32_t m = __ballot_sync(FULL, in);
                const uint32_t l0 = in ? cl : 0u, h0 = in ? ch : 0u;
                const u64 sr = (u64)__reduce_add_sync(FULL, l0 * r) + ((u64)__reduce_add_sync(FULL, h0 * r) << 16);
                const u64 sg = (u64)__reduce_add_sync(FULL, l0 * g) + ((u64)__reduce_add_sync(FULL, h0 * g) << 16);
                const u64 sb = (u64)__reduce_add_sync(FULL, l0 * b) + ((u64)__reduce_add_sync(FULL, h0 * b) << 16);
                const u64 sn = (u64)__reduce_add_sync(FULL, l0) + ((u64)__reduce_add_sync(FULL, h0) << 16);
                if (lane == 0) {
                    s_acc[warp][k][0] += sr;
                    s_acc[warp][k][1] += sg;
                    s_acc[warp][k][2] += sb;
                    s_acc[warp][k][3] += sn;
                }
                todo &= ~m;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < K * 4; i += HT) {
        u64 t = 0;
        for (int w = 0; w < HT / 32; ++w) t += s_acc[w][i >> 2][i & 3];
        if (t) atomicAdd(&sums[i], t);
    }
}

__global__ void __launch_bounds__(HT) k_hist_step(const uint32_t* __restrict__ keys,
                                                  const uint32_t* __restrict__ counts, size_t n, int K,
                                                  const float* __restrict__ centers, u64* sums,
                                                  uint8_t* __restrict__ labels_out,
                                                  const int32_t* __restrict__ state, const int32_t* __restrict__ n_dev) {
    if (state && (state[1] | state[3])) return;
    if (n_dev) {   // the list length lives on the device (no host round trip after the compaction)
        const size_t nd = (size_t)max(*n_dev, 0);
        n = nd < n ? nd : n;
        if ((size_t)blockIdx.x * (HT * EPS_T) >= n) return;
    }
    __shared__ u64 s_nc[KMAX][3];   // (-c, -c)
    __shared__ u64 s_acc[HT / 32][KMAX][4];
    hist_step_body(keys, counts, n, K, centers, sums, labels_out, s_nc, s_acc);
}

// ---- the colour table across ranks without NCCL: barrier, then every rank PULLS its share from the peers ---------------
__global__ void __launch_bounds__(128) k_p2p_barrier(P2PMailbox* const* peers, int rank, int world) {
    p2p_barrier(peers, rank, world);
}

// share[j * HB + i] = sum over ranks q of table_q[(j * world + rank) * HB + i]: the interleaved 2048-key blocks this rank
// owns, read straight out of every peer's HBM over NVLink (16-byte loads), summed, packed back to back -- what the
// reduce-scatter of the block-transposed table delivered, without the transposition copy and the NCCL launch.
__global__ void __launch_bounds__(256) k_hist_pull_reduce(const uint32_t* const* tables, int rank, int world,
                                                          uint32_t* __restrict__ share) {
    const size_t n4 = (size_t)(1u << 24) / world / 4;       // uint4 elements of the share
    for (size_t v = (size_t)blockIdx.x * 256 + threadIdx.x; v < n4; v += (size_t)gridDim.x * 256) {
        const size_t j = v / (HB / 4), i4 = v % (HB / 4);
        const size_t src = (j * world + rank) * (HB / 4) + i4;
        uint4 acc = make_uint4(0u, 0u, 0u, 0u);
        for (int q = 0; q < world; ++q) {
            const uint4 a = __ldcg(reinterpret_cast<const uint4*>(tables[q]) + src);
            acc.x += a.x;
            acc.y += a.y;
            acc.z += a.z;
            acc.w += a.w;
        }
        reinterpret_cast<uint4*>(share)[v] = acc;
    }
}

// The whole Lloyd loop of the colour-histogram form as ONE persistent cooperative kernel per rank: assignment over this
// rank's (key, count) entries by every CTA, grid barrier, exchange of the K x 4 partial sums with the other ranks through
// their NVLink mailboxes + centre update by CTA 0 (k_kmeans_p2p.cuh), grid barrier, next iteration -- until the state says
// converged (every rank sees the same state) or frozen (empty cluster: the host repairs and launches again) or `iters`
// iterations have run.  No launch, no host look and no NCCL call between iterations; the entries (a few thousand for a
// design, <= 2^24) and the centres stay in L2.
__global__ void __launch_bounds__(HT) k_hist_lloyd(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ counts,
                                                   size_t n, const int32_t* __restrict__ n_dev, int K, float* centers,
                                                   u64* partial, uint8_t* __restrict__ labels_out, P2PMailbox* const* peers,
                                                   int rank, int world, int max_iter, double eps2, int32_t* state,
                                                   double* shift_out, u64* totals, int iters) {
    cg::grid_group grid = cg::this_grid();
    __shared__ u64 s_nc[KMAX][3];
    __shared__ u64 s_acc[HT / 32][KMAX][4];
    __shared__ u64 s_tot[KMAX * 4];
    if (n_dev) {
        const size_t nd = (size_t)max(*n_dev, 0);
        n = nd < n ? nd : n;
    }
    for (int it = 0; it < iters; ++it) {
        const volatile int32_t* vs = state;
        if (vs[1] | vs[3]) break;                 // grid-uniform: written before the last grid barrier
        if ((size_t)blockIdx.x * (HT * EPS_T) < n)
            hist_step_body(keys, counts, n, K, centers, partial, labels_out, s_nc, s_acc);
        __threadfence();
        grid.sync();
        if (blockIdx.x == 0)
            p2p_exchange_and_update(K, partial, peers, rank, world, centers, max_iter, eps2, state, shift_out, totals, s_tot);
        __threadfence();
        grid.sync();
    }
}

// largest float32 distance to `base` among the colours assigned to `donor` (bits of a non-negative float
// order like the float): the threshold that lets the pixel pass below skip almost every pixel
__global__ void __launch_bounds__(HT) k_hist_farthest(const uint32_t* __restrict__ keys, size_t n, int K,
                                                      const float* __restrict__ centers, int donor, float b0, float b1,
                                                      float b2, uint32_t* out) {
    __shared__ float s_c[KMAX][3];
    for (int i = threadIdx.x; i < K * 3; i += HT) (&s_c[0][0])[i] = centers[i];
    __syncthreads();
    const float base[3] = {b0, b1, b2};
    uint32_t best = 0u;
    const size_t stride = (size_t)gridDim.x * HT;
    for (size_t i = blockIdx.x * (size_t)HT + threadIdx.x; i < n; i += stride) {
        const uint32_t key = keys[i];
        const float fr = (float)(key >> 16), fg = (float)((key >> 8) & 0xffu), fb = (float)(key & 0xffu);
        float bd = fdist3(fr, fg, fb, s_c[0]);
        int bl = 0;
        for (int k = 1; k < K; ++k) {
            const float d = fdist3(fr, fg, fb, s_c[k]);
            if (d < bd) {
                bd = d;
                bl = k;
            }
        }
        if (bl != donor) continue;
        best = max(best, __float_as_uint(fdist3(fr, fg, fb, base)));
    }
    best = __reduce_max_sync(FULL, best);
    if ((threadIdx.x & 31) == 0 && best) atomicMax(out, best);
}

struct SkipList {
    int n;
    uint32_t idx[KMAX];
};

// farthest member (f32 distance to `base`) of cluster `donor` under the assignment to `centers`; result = max
// over pixels of (dist bits << 32 | (pixel index + index_base)) + 1.  With want_bits != 0 only pixels at exactly
// that float32 distance from `base` are looked at (the answer's distance is known from the distinct colours).
__device__ __forceinline__ void farthest_px(uint32_t r, uint32_t g, uint32_t b, size_t gidx, const float (*s_c)[3], int K,
                                            int donor, const float* base, uint32_t want_bits, const SkipList& skip,
                                            u64& best) {
    const float fr = (float)r, fg = (float)g, fb = (float)b;
    const uint32_t db = __float_as_uint(fdist3(fr, fg, fb, base));
    if (want_bits && db != want_bits) return;
    float bd = fdist3(fr, fg, fb, s_c[0]);
    int bl = 0;
    for (int k = 1; k < K; ++k) {
        const float d = fdist3(fr, fg, fb, s_c[k]);
        if (d < bd) {
            bd = d;
            bl = k;
        }
    }
    if (bl != donor) return;
    for (int j = 0; j < skip.n; ++j)   // pixels an earlier repair of this update already moved out of the donor
        if (skip.idx[j] == (uint32_t)gidx) return;
    const u64 cand = (((u64)db << 32) | (uint32_t)gidx) + 1ull;
    best = cand > best ? cand : best;
}

__global__ void __launch_bounds__(HT) k_pixels_farthest(const uint8_t* __restrict__ bgr, size_t npix, int head, int K,
                                                        const float* __restrict__ centers, int donor, float b0,
                                                        float b1, float b2, uint32_t index_base, SkipList skip,
                                                        uint32_t want_bits, u64* out) {
    __shared__ float s_c[KMAX][3];
    for (int i = threadIdx.x; i < K * 3; i += HT) (&s_c[0][0])[i] = centers[i];
    __syncthreads();
    const float base[3] = {b0, b1, b2};
    u64 best = 0ull;
    const size_t nbulk = npix > (size_t)head ? (npix - head) / 16 : 0;
    const uint4* src = reinterpret_cast<const uint4*>(bgr + (size_t)head * 3);
    const size_t gstride = (size_t)gridDim.x * HT;
    for (size_t g = blockIdx.x * (size_t)HT + threadIdx.x; g < nbulk; g += gstride) {
        uint32_t w[13], key[16];
        load48(src + 3 * g, w);
        keys16(w, key);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            farthest_px(key[j] >> 16, (key[j] >> 8) & 0xffu, key[j] & 0xffu, (size_t)head + 16 * g + j + index_base, s_c,
                        K, donor, base, want_bits, skip, best);
    }
    if (blockIdx.x == 0) {
        const size_t hd = npix < (size_t)head ? npix : (size_t)head;
        const size_t tail0 = (size_t)head + nbulk * 16;
        const size_t nrest = hd + (npix > tail0 ? npix - tail0 : 0);
        for (size_t q = threadIdx.x; q < nrest; q += HT) {
            const size_t p = q < hd ? q : tail0 + (q - hd);
            farthest_px(bgr[3 * p + 2], bgr[3 * p + 1], bgr[3 * p], p + index_base, s_c, K, donor, base, want_bits, skip,
                        best);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u64 nb = __shfl_xor_sync(FULL, best, o);
        best = nb > best ? nb : best;
    }
    if ((threadIdx.x & 31) == 0 && best) atomicMax(out, best);
}

__global__ void __launch_bounds__(HT) k_hist_lut(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ labels,
                                                 size_t n, uint8_t* __restrict__ lut) {
    const size_t stride = (size_t)gridDim.x * HT;
    for (size_t i = blockIdx.x * (size_t)HT + threadIdx.x; i < n; i += stride) lut[keys[i]] = labels[i];
}

// per-pixel labels through the colour -> label table
__global__ void __launch_bounds__(HT) k_pixels_lookup(const uint8_t* __restrict__ bgr, size_t npix, int head,
                                                      const uint8_t* __restrict__ lut, uint8_t* __restrict__ labels) {
    const size_t nbulk = npix > (size_t)head ? (npix - head) / 16 : 0;
    const uint4* base = reinterpret_cast<const uint4*>(bgr + (size_t)head * 3);
    const size_t gstride = (size_t)gridDim.x * HT;
    for (size_t g = blockIdx.x * (size_t)HT + threadIdx.x; g < nbulk; g += gstride) {
        uint32_t w[13], key[16], o[4] = {0u, 0u, 0u, 0u};
        load48(base + 3 * g, w);
        keys16(w, key);
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j >> 2] |= (uint32_t)__ldg(lut + key[j]) << (8 * (j & 3));
        uint8_t* dst = labels + head + 16 * g;
        if (((uintptr_t)dst & 15) == 0) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[j] = (uint8_t)(o[j >> 2] >> (8 * (j & 3)));
        }
    }
    if (blockIdx.x == 0) {
        const size_t hd = npix < (size_t)head ? npix : (size_t)head;
        const size_t tail0 = (size_t)head + nbulk * 16;
        const size_t nrest = hd + (npix > tail0 ? npix - tail0 : 0);
        for (size_t q = threadIdx.x; q < nrest; q += HT) {
            const size_t p = q < hd ? q : tail0 + (q - hd);
            labels[p] = lut[((uint32_t)bgr[3 * p + 2] << 16) | ((uint32_t)bgr[3 * p + 1] << 8) | bgr[3 * p]];
        }
    }
}

// pixels before the first 16-byte boundary that is also a pixel boundary: 3 h = -addr (mod 16)
inline int head_pixels(const void* p) { return (int)(((16 - ((uintptr_t)p & 15)) & 15) * 11 % 16); }

inline unsigned stream_grid(const llfe_ctx* ctx, size_t items_per_thread_groups) {
    const size_t want = ceil_div_sz(items_per_thread_groups, HT);
    const size_t cap = (size_t)ctx->sm_count * 16;
    return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

extern "C" int llfe_pixels_histogram(llfe_ctx* ctx, const uint8_t* d_bgr, size_t n_pixels, uint32_t* d_hist) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_hist != nullptr && (d_bgr != nullptr || n_pixels == 0));
    LLFE_CHECK_ARG(n_pixels <= 0x7fffffffull);   // a bin must stay below 2^31 (ranks all-reduce the table as int32)
    if (n_pixels == 0) return LLFE_OK;
    LLFE_KERNEL(ctx, "k_hist_count");
    k_hist_count<<<stream_grid(ctx, n_pixels / 16 + 1), HT, 0, ctx->stream>>>(d_bgr, n_pixels, head_pixels(d_bgr), d_hist);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

extern "C" int llfe_histogram_compact(llfe_ctx* ctx, const uint32_t* d_hist, int part, int parts, uint32_t* d_keys_or_null,
                                      uint32_t* d_counts_or_null, size_t cap, int32_t* d_n, int packed) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_hist != nullptr && d_n != nullptr && parts >= 1 && part >= 0 && part < parts);
    LLFE_CHECK_ARG(!packed || NBLK % parts == 0);
    LLFE_CHECK_ARG(cap == 0 || (d_keys_or_null != nullptr && d_counts_or_null != nullptr));
    void* ws = nullptr;
    LLFE_TRY(llfe_workspace(ctx, 2 * WsCarver::need(NBLK * sizeof(uint32_t)), &ws));
    WsCarver carve(ws);
    uint32_t* bcount = carve.take<uint32_t>(NBLK);
    uint32_t* boffs = carve.take<uint32_t>(NBLK);
    LLFE_KERNEL(ctx, "k_hist_blockcount");
    k_hist_blockcount<<<NBLK, HT, 0, ctx->stream>>>(d_hist, part, parts, packed, bcount);
    LLFE_LAUNCHED(ctx);
    LLFE_KERNEL(ctx, "k_hist_blockscan");
    k_hist_blockscan<<<1, 1024, 0, ctx->stream>>>(bcount, boffs, d_n);
    LLFE_LAUNCHED(ctx);
    if (cap > 0) {
        LLFE_KERNEL(ctx, "k_hist_emit");
        k_hist_emit<<<NBLK, HT, 0, ctx->stream>>>(d_hist, part, parts, packed, boffs, d_keys_or_null, d_counts_or_null, cap);
        LLFE_LAUNCHED(ctx);
    }
    return LLFE_OK;
}

extern "C" int llfe_kmeans_hist_step(llfe_ctx* ctx, const uint32_t* d_keys, const uint32_t* d_counts, size_t n, int k,
                                     const float* d_centers, uint64_t* d_sums_counts, uint8_t* d_labels_or_null,
                                     const int32_t* d_state_or_null, const int32_t* d_n_or_null) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_centers != nullptr && d_sums_counts != nullptr && k >= 1 && k <= KMAX);
    LLFE_CHECK_ARG(n == 0 || (d_keys != nullptr && d_counts != nullptr));
    if (n == 0) return LLFE_OK;
    const size_t want = ceil_div_sz(n, HT * EPS_T);
    const size_t cap = (size_t)ctx->sm_count * 8;
    LLFE_KERNEL(ctx, "k_hist_step");
    k_hist_step<<<(unsigned)(want > cap ? cap : want), HT, 0, ctx->stream>>>(
        d_keys, d_counts, n, k, d_centers, (u64*)d_sums_counts, d_labels_or_null, d_state_or_null, d_n_or_null);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

extern "C" int llfe_p2p_barrier(llfe_ctx* ctx, void* const* d_mailboxes, int rank, int world) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_mailboxes != nullptr && world >= 1 && world <= P2P_MAXW && rank >= 0 && rank < world);
    LLFE_KERNEL(ctx, "k_p2p_barrier");
    k_p2p_barrier<<<1, 128, 0, ctx->stream>>>((P2PMailbox* const*)d_mailboxes, rank, world);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

extern "C" int llfe_histogram_pull_reduce(llfe_ctx* ctx, const uint32_t* const* d_tables, int rank, int world,
                                          uint32_t* d_share) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_tables != nullptr && d_share != nullptr && world >= 1 && world <= P2P_MAXW && rank >= 0 && rank < world);
    LLFE_CHECK_ARG(NBLK % world == 0);
    LLFE_KERNEL(ctx, "k_hist_pull_reduce");
    k_hist_pull_reduce<<<stream_grid(ctx, (size_t)(1u << 24) / world / 4), 256, 0, ctx->stream>>>(d_tables, rank, world, d_share);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

extern "C" int llfe_kmeans_hist_lloyd(llfe_ctx* ctx, const uint32_t* d_keys, const uint32_t* d_counts, size_t n,
                                      const int32_t* d_n_or_null, int k, float* d_centers, uint64_t* d_partial_sums,
                                      uint8_t* d_labels_or_null, void* const* d_mailboxes_or_null, int rank, int world,
                                      int max_iter, double eps, int32_t* d_state, double* d_shift, uint64_t* d_totals,
                                      int iterations) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_centers != nullptr && d_partial_sums != nullptr && d_state != nullptr && d_totals != nullptr &&
                   d_totals != d_partial_sums && k >= 1 && k <= KMAX && max_iter >= 1 && iterations >= 1);
    LLFE_CHECK_ARG(n == 0 || (d_keys != nullptr && d_counts != nullptr));
    LLFE_CHECK_ARG(world >= 1 && world <= P2P_MAXW && rank >= 0 && rank < world && (world == 1 || d_mailboxes_or_null != nullptr));
    int per_sm = 0;
    LLFE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_hist_lloyd, HT, 0));
    LLFE_CHECK_ARG(per_sm >= 1);
    const size_t want = n ? ceil_div_sz(n, HT * EPS_T) : 1;
    const size_t cap = (size_t)ctx->sm_count * (per_sm < 8 ? per_sm : 8);
    const unsigned grid = (unsigned)(want > cap ? cap : want);
    const double eps2 = eps * eps;
    P2PMailbox* const* peers = (P2PMailbox* const*)d_mailboxes_or_null;
    u64* partial = (u64*)d_partial_sums;
    u64* totals = (u64*)d_totals;
    void* args[] = {&d_keys, &d_counts, &n, &d_n_or_null, &k, &d_centers, &partial, &d_labels_or_null, &peers, &rank, &world,
                    &max_iter, (void*)&eps2, &d_state, &d_shift, &totals, &iterations};
    LLFE_KERNEL(ctx, "k_hist_lloyd");
    LLFE_CUDA(cudaLaunchCooperativeKernel((const void*)k_hist_lloyd, dim3(grid), dim3(HT), args, 0, ctx->stream));
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

extern "C" int llfe_hist_labels_to_lut(llfe_ctx* ctx, const uint32_t* d_keys, const uint8_t* d_labels, size_t n,
                                       uint8_t* d_lut) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_lut != nullptr && (n == 0 || (d_keys != nullptr && d_labels != nullptr)));
    if (n == 0) return LLFE_OK;
    LLFE_KERNEL(ctx, "k_hist_lut");
    k_hist_lut<<<stream_grid(ctx, n), HT, 0, ctx->stream>>>(d_keys, d_labels, n, d_lut);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

extern "C" int llfe_pixels_lookup(llfe_ctx* ctx, const uint8_t* d_bgr, size_t n_pixels, const uint8_t* d_lut,
                                  uint8_t* d_labels) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_lut != nullptr && (n_pixels == 0 || (d_bgr != nullptr && d_labels != nullptr)));
    if (n_pixels == 0) return LLFE_OK;
    LLFE_KERNEL(ctx, "k_pixels_lookup");
    k_pixels_lookup<<<stream_grid(ctx, n_pixels / 16 + 1), HT, 0, ctx->stream>>>(d_bgr, n_pixels, head_pixels(d_bgr), d_lut,
                                                                                 d_labels);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

extern "C" int llfe_kmeans_hist_farthest(llfe_ctx* ctx, const uint32_t* d_keys, size_t n, int k, const float* d_centers,
                                         int donor, const float* h_base3, uint32_t* d_out_bits) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_centers != nullptr && h_base3 != nullptr && d_out_bits != nullptr);
    LLFE_CHECK_ARG(k >= 1 && k <= KMAX && donor >= 0 && donor < k && (n == 0 || d_keys != nullptr));
    if (n == 0) return LLFE_OK;
    LLFE_KERNEL(ctx, "k_hist_farthest");
    k_hist_farthest<<<stream_grid(ctx, n), HT, 0, ctx->stream>>>(d_keys, n, k, d_centers, donor, h_base3[0], h_base3[1],
                                                                 h_base3[2], d_out_bits);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

extern "C" int llfe_kmeans_pixels_farthest(llfe_ctx* ctx, const uint8_t* d_bgr, size_t n_pixels, int k,
                                           const float* d_centers, int donor, const float* h_base3,
                                           uint32_t index_base, const uint32_t* h_skip, int n_skip, uint32_t want_dist_bits,
                                           uint64_t* d_out) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_centers != nullptr && h_base3 != nullptr && d_out != nullptr);
    LLFE_CHECK_ARG(n_pixels == 0 || d_bgr != nullptr);
    LLFE_CHECK_ARG(n_skip >= 0 && n_skip <= KMAX && (n_skip == 0 || h_skip != nullptr));
    SkipList skip;
    skip.n = n_skip;
    for (int j = 0; j < n_skip; ++j) skip.idx[j] = h_skip[j];
    LLFE_CHECK_ARG(k >= 1 && k <= KMAX && donor >= 0 && donor < k && n_pixels + index_base <= 0xffffffffull);
    if (n_pixels == 0) return LLFE_OK;
    LLFE_KERNEL(ctx, "k_pixels_farthest");
    k_pixels_farthest<<<stream_grid(ctx, n_pixels / 16 + 1), HT, 0, ctx->stream>>>(
        d_bgr, n_pixels, head_pixels(d_bgr), k, d_centers, donor, h_base3[0], h_base3[1], h_base3[2], index_base, skip,
        want_dist_bits, (u64*)d_out);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}
