// Fast path of cv2.kmeans over UNWEIGHTED colour lists (the reference's palette call,
// color_extractor.py:189-196): same arithmetic and results as k_kmeans (k_kmeans.cu), laid
// out for the SM:
//   * one CTA per (attempt, image), TWO CTAs (32 warps) per SM: the kernel is held to 64 registers
//     and keeps ONE 32-bit state word per colour in shared memory for the whole solve (the kmeans++
//     distance during the seeding; label | upper bound | lower bound during Lloyd).  The colour keys
//     themselves are re-read from the (L2-resident, read-only) sorted key list when a pass needs
//     them, so no iteration touches HBM.  Lists too long for shared memory run the same code on
//     a global scratch copy (second instantiation, IN_SMEM = false);
//   * warp-blocked point ownership (warp w owns a contiguous block, lanes stride it): bank-
//     conflict-free and contiguous for the kmeans++ prefix search;
//   * kmeans++ (cv::generateCentersPP): integer squared distances in two instructions
//     (per-byte |a-b| + IDP.4A); the three trial centres of a step are drawn first (the draws
//     do not depend on the trials) and evaluated in ONE pass over the points; the sequential
//     "p -= d; if (p <= 0) break" walk is an exact prefix search (warp totals -> row totals ->
//     one warp scan);
//   * Lloyd: the first assignment is cv2's exact float32 search for every point with packed
//     16-bit register accumulators + REDUX; later assignments use Hamerly bounds (see below) so
//     that only points that can change their label pay for the search, with exact integer sums
//     updated incrementally.
#include <stdlib.h>

#include <type_traits>

#include "llfe_common.cuh"
#include "llfe_device.cuh"
#include "k_kmeans_shared.cuh"

namespace {

constexpr int FT_STD = 512;    // two CTAs per SM (lists of up to pts2 colours) and the global-scratch variant
constexpr int FT_LONG = 1024;  // one CTA per SM with all of its shared memory: twice the warps keep the SM as busy
constexpr int QROWS = 8;    // rows (of 32 points) per bounds-test / drain block
// static shared memory (centres, totals, queues, row totals, reduction scratch), upper bounds per variant
constexpr size_t STATIC_SMEM_BOUND = 18 * 1024, STATIC_SMEM_BOUND_LONG = 25 * 1024;
constexpr int RMAX_POINTS = 112 * 32 * 16;   // most points a shared-memory list may have (row-total table size)
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float fdist4(float r, float g, float b, float4 c) {
    float t0 = __fsub_rn(r, c.x), t1 = __fsub_rn(g, c.y), t2 = __fsub_rn(b, c.z);
    float d = __fmul_rn(t0, t0);
    d = __fadd_rn(d, __fmul_rn(t1, t1));
    d = __fadd_rn(d, __fmul_rn(t2, t2));
    return d;
}

// Square root for the BOUNDS only (never for a value cv2 computes): sqrt.approx is within 2^-23 relative, i.e. 0.002 of
// the 1/32-unit grid at the largest RGB distance, and q_up / q_dn each keep a whole unit of slack on top of their rounding.
__device__ __forceinline__ float bound_sqrt(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// State word of a point during Lloyd: label | ub | lb.  K <= 16: 4 + 14 + 14 bits, bounds in 1/32 colour
// units (the largest RGB distance, 441.7, is 14134 units); K <= 32: 5 + 13 + 14 bits in 1/16 units.
// q_up rounds up, q_dn down, each with one extra unit of slack.
template <int KC>
struct Fmt {
    static constexpr int LB = KC <= 16 ? 4 : 5;            // label bits
    static constexpr int UB = KC <= 16 ? 14 : 13;          // upper-bound bits
    static constexpr uint32_t UMAX = (1u << UB) - 1u;      // "unknown": never passes a bound test
    static constexpr uint32_t LMAX = (1u << 14) - 1u;
    static constexpr float SCALE = KC <= 16 ? 32.f : 16.f;
    __device__ static __forceinline__ uint32_t q_up(float d) { return min(UMAX, (uint32_t)(d * SCALE) + 2u); }
    __device__ static __forceinline__ uint32_t q_dn(float d) {
        const float v = fminf(d * SCALE, (float)LMAX);
        return v >= 2.f ? (uint32_t)v - 1u : 0u;
    }
    __device__ static __forceinline__ uint32_t pack(uint32_t a, uint32_t ub, uint32_t lb) {
        return (a << (32 - LB)) | (ub << 14) | lb;
    }
    __device__ static __forceinline__ uint32_t label(uint32_t x) { return x >> (32 - LB); }
    __device__ static __forceinline__ uint32_t ub(uint32_t x) { return (x >> 14) & UMAX; }
    __device__ static __forceinline__ uint32_t lb(uint32_t x) { return x & LMAX; }
};

// bytes 2/1/0 of the key as exact floats without the conversion pipe: 2^23 + byte, minus 2^23
__device__ __forceinline__ void unpackf(uint32_t key, float& r, float& g, float& b) {
    r = __uint_as_float(__byte_perm(key, 0x4B000000u, 0x7652)) - 8388608.0f;
    g = __uint_as_float(__byte_perm(key, 0x4B000000u, 0x7651)) - 8388608.0f;
    b = __uint_as_float(__byte_perm(key, 0x4B000000u, 0x7650)) - 8388608.0f;
}

// squared RGB distance of two keys whose top bytes are equal: per-byte |a-b|, then sum of squares
__device__ __forceinline__ uint32_t idist2(uint32_t a, uint32_t b) {
    const uint32_t d = __vabsdiffu4(a, b);
    return __dp4a(d, d, 0u);
}


// ---- cv2's sequential float32 centre sums beyond 2^24 -------------------------------------------------------------
// cv::kmeans accumulates `center[j] += sample[j]` in float32, point after point in index order.  Below 2^24 every
// partial sum is an exact integer; above, each addition rounds to the float32 grid (ulp u = 2, 4, ...), with ties to
// even -- so the result depends on the order and differs from the exact integer sum.  Lists long enough for that
// (> 65 793 members in one cluster: photo-like frames, SURVEY 8(d)'s U ~ 1.95 M set) only run in the global-scratch
// variant of the kernel, which reproduces the sequential sum operation for operation, in parallel:
//   * inside one binade (fixed u = 2^m) an addition maps the running sum s to s + u * inc(x, parity(s / u)), so an
//     element is a function {parity 0, parity 1} -> increment, and functions compose associatively: every thread
//     folds its contiguous segment of the list for both incoming parities, an ordered scan composes the segments,
//     and every thread then knows the exact sum entering its segment;
//   * a second walk finds the first element whose exact sum t = s + x reaches the next binade; that one addition is
//     done with a real float add, and the next phase continues behind it with u doubled (at most 6 phases up to 2^29).
// The three channels of a cluster are folded in the same walks.  Verified against np.cumsum(float32) on the host
// (oracle/cvops.py:_sums_f32_sequential is that call) and against cv2.kmeans itself (tests/test_gpu_long_lists.py).
__device__ __forceinline__ uint32_t seq_inc(uint32_t x, uint32_t parity, int m) {
    if (m == 0) return x;
    const uint32_t u = 1u << m, half = u >> 1, q = x >> m, r = x & (u - 1u);
    return q + ((r > half || (r == half && ((parity + q) & 1u))) ? 1u : 0u);
}

struct SeqShared {
    uint32_t comp[32][6];      // per-warp composites: [channel][start parity]
    uint32_t cur[3];           // running float32 sum per channel (an integer that float32 represents exactly)
    int start[3];              // next list index of the channel's chain
    int done[3];
    int cross_i[3];
    uint32_t total[6];
};

template <typename F, int FT>
__device__ void seq_f32_sums3(int k, int U, const uint32_t* __restrict__ aux, const uint32_t* __restrict__ keys,
                              SeqShared& S, float* out3) {
    constexpr int FW = FT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seg = (U + FT - 1) / FT, lo = min(U, tid * seg), hi = min(U, lo + seg);
    if (tid < 3) {
        S.cur[tid] = 0u;
        S.start[tid] = 0;
        S.done[tid] = 0;
    }
    __syncthreads();
    for (;;) {
        uint32_t s[3], limit[3], par[3];
        int st[3], m[3];
        bool act[3], any = false;
        int first = U;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s[c] = S.cur[c];
            st[c] = S.start[c];
            act[c] = !S.done[c];
            m[c] = s[c] < (1u << 24) ? 0 : (31 - __clz(s[c])) - 23;
            limit[c] = 1u << (24 + m[c]);
            par[c] = (s[c] >> m[c]) & 1u;
            any |= act[c];
            if (act[c]) first = min(first, st[c]);
        }
        if (!any) break;   // block-uniform (shared state)
        __syncthreads();   // everyone has read the state before it is rewritten
        if (tid < 3) S.cross_i[tid] = 0x7fffffff;
        // ---- walk 1: this thread's segment as a function of the incoming parity, per channel
        uint32_t d[3][2] = {{0u, 0u}, {0u, 0u}, {0u, 0u}};
        for (int i = max(lo, first); i < hi; ++i) {
            if ((int)F::label(aux[i]) != k) continue;
            const uint32_t key = __ldg(keys + i);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (!act[c] || i < st[c]) continue;
                const uint32_t x = (key >> (16 - 8 * c)) & 255u;
                d[c][0] += seq_inc(x, d[c][0] & 1u, m[c]);
                d[c][1] += seq_inc(x, (1u + d[c][1]) & 1u, m[c]);
            }
        }
        // ---- ordered scan of the composites: inclusive inside the warp, then the warps in order
        uint32_t inc0[3], inc1[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            inc0[c] = d[c][0];
            inc1[c] = d[c][1];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t e0 = __shfl_up_sync(FULL, inc0[c], o), e1 = __shfl_up_sync(FULL, inc1[c], o);
                if (lane >= o) {   // (earlier segment E) then (mine): C_p = E_p + mine_{p + E_p}
                    const uint32_t n0 = e0 + ((e0 & 1u) ? inc1[c] : inc0[c]);
                    const uint32_t n1 = e1 + (((1u + e1) & 1u) ? inc1[c] : inc0[c]);
                    inc0[c] = n0;
                    inc1[c] = n1;
                }
            }
            if (lane == 31) {
                S.comp[warp][2 * c] = inc0[c];
                S.comp[warp][2 * c + 1] = inc1[c];
            }
        }
        __syncthreads();
        uint32_t sin[3];   // exact float32 running sum entering this thread's segment
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint32_t acc = 0u;
            for (int w = 0; w < warp; ++w) acc += S.comp[w][2 * c + ((par[c] + acc) & 1u)];
            const uint32_t x0 = __shfl_up_sync(FULL, inc0[c], 1), x1 = __shfl_up_sync(FULL, inc1[c], 1);
            if (lane > 0) acc += ((par[c] + acc) & 1u) ? x1 : x0;
            sin[c] = s[c] + (acc << m[c]);
            if (tid == FT - 1) S.total[c] = acc + (((par[c] + acc) & 1u) ? d[c][1] : d[c][0]);
        }
        // ---- walk 2: the first element whose exact sum reaches the next binade
        int my_i[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
        uint32_t my_s[3] = {0u, 0u, 0u}, my_x[3] = {0u, 0u, 0u};
        for (int i = max(lo, first); i < hi; ++i) {
            if ((int)F::label(aux[i]) != k) continue;
            const uint32_t key = __ldg(keys + i);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (!act[c] || i < st[c] || my_i[c] != 0x7fffffff) continue;
                const uint32_t x = (key >> (16 - 8 * c)) & 255u;
                if (sin[c] + x >= limit[c]) {
                    my_i[c] = i;
                    my_s[c] = sin[c];
                    my_x[c] = x;
                    atomicMin(&S.cross_i[c], i);
                } else {
                    sin[c] += seq_inc(x, (sin[c] >> m[c]) & 1u, m[c]) << m[c];
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (!act[c]) continue;
            const int ci = S.cross_i[c];
            if (ci == 0x7fffffff) {
                if (tid == FT - 1) {
                    S.cur[c] = s[c] + (S.total[c] << m[c]);
                    S.done[c] = 1;
                }
            } else if (my_i[c] == ci) {
                S.cur[c] = (uint32_t)__fadd_rn((float)my_s[c], (float)my_x[c]);   // the addition that changes the binade
                S.start[c] = ci + 1;
            }
        }
        __syncthreads();
    }
    if (tid < 3) out3[tid] = (float)S.cur[tid];
    __syncthreads();
}

template <int KC, bool IN_SMEM, int FT>
__global__ void __launch_bounds__(FT, FT == FT_LONG ? 1 : 2) k_kmeans_fast(KmParams P, int u_lo, int smem_points, int img_base) {
    typedef Fmt<KC> F;
    constexpr int FW = FT / 32;
    constexpr int RMAX = RMAX_POINTS / (32 * FW);   // rows per warp the row-total table holds
    extern __shared__ uint32_t dyn[];
    const int att = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int img = (IN_SMEM && P.order) ? P.order[blockIdx.y] : blockIdx.y + img_base;
    const int U = min(P.count[img], P.max_unique);
    const int K = min(P.k, U);
    // this launch owns the images with u_lo < U <= smem_points (IN_SMEM) / U > smem_points (global scratch)
    if (IN_SMEM ? (U <= u_lo || U > smem_points) : (U <= smem_points)) return;
    const size_t slot = (size_t)img * P.attempts + att;
    if (K <= 1 || P.count[img] > P.max_unique) {   // nothing to cluster / list truncated (see k_kmeans_pick)
        if (tid == 0) {
            P.compact[slot] = 0.0;
            P.iters[slot] = 0;
            P.inexact[slot] = 0;
        }
        return;
    }
    const uint32_t* __restrict__ keys = P.keys + (size_t)img * P.max_unique;
    // per-point state word: shared memory when the list fits, else this slot's global scratch
    // (the global scratch is sized for one launch of `dist_images` images: index it by blockIdx.y)
    uint32_t* const aux = IN_SMEM ? dyn : P.dist + ((size_t)blockIdx.y * P.attempts + att) * 2 * (size_t)P.max_unique;
    auto key_at = [&](int i) -> uint32_t { return __ldg(keys + i) & 0xffffffu; };
    // warp-blocked ownership: rows of 32 consecutive points, `rows` rows per warp
    const int rows = (U + 32 * FW - 1) / (32 * FW);
    const int wbase = warp * rows * 32;

    __shared__ float4 s_c[KMAX];
    __shared__ float4 s_old[KMAX];
    __shared__ float4 s_asg[KMAX];          // the centres the current labels were assigned with
    __shared__ int s_sum[KMAX][4];          // exact per-cluster {sum R, sum G, sum B, count}
    __shared__ float s_fsum[KMAX][3];       // the float32 sums cv2 has: == s_sum below 2^24, sequential rounding above
    // scratch of seq_f32_sums3: only the global-scratch variant gets there, the others keep their shared memory
    __shared__ typename std::conditional<IN_SMEM, int, SeqShared>::type s_seq;
    __shared__ uint32_t s_dq[KMAX];
    __shared__ uint4 s_tab[KMAX];           // per label: {own centre's drift, largest drift of another centre, half gap, -}
    __shared__ uint8_t s_queue[FW][QROWS * 32];   // per-warp queue of points whose bounds failed
    // per-warp ring of points that need the exact search (row of the warp << 5 | lane: 12 bits for lists in shared
    // memory, more for the global-scratch variant)
    typedef typename std::conditional<IN_SMEM, uint16_t, uint32_t>::type Q2;
    __shared__ Q2 s_queue2[FW][64];
    __shared__ uint32_t s_rowsum[IN_SMEM ? FW : 1][IN_SMEM ? RMAX : 1];
    __shared__ unsigned long long s_wtot[FW];
    __shared__ unsigned long long s_red3[FW][3];
    __shared__ double s_redd[FW];
    __shared__ int s_ci[3];
    __shared__ double s_p[3];
    __shared__ int s_flag;
    __shared__ unsigned int s_cnt[4];
    __shared__ unsigned long long s_far;

    const long long t_start = clock64();
    long long t_first = 0;
    if (tid == 0) s_flag = 0;
    if (tid < 4) s_cnt[tid] = 0;
    __syncthreads();
    const long long t_loaded = clock64();

    int it;
    if (P.init) {
        if (tid < KMAX) {
            float4 c = make_float4(3e18f, 3e18f, 3e18f, 0.f);
            if (tid < K) {
                const float* p = P.init + ((size_t)img * P.k + tid) * 3;
                c = make_float4(p[0], p[1], p[2], 0.f);
            }
            s_c[tid] = c;
            s_old[tid] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        it = 0;
        __syncthreads();
    } else {
        // ---------------- kmeans++ (cv::generateCentersPP, 3 trials per centre) ----------------
        unsigned long long rs = 0;
        if (tid < KMAX) s_c[tid] = make_float4(3e18f, 3e18f, 3e18f, 0.f);
        if (tid == 0) {
            rs = P.rng_state[img];
            if (rs == 0) rs = 0xffffffffull;
            const int per_attempt = 1 + 6 * (K - 1);
            for (int i = 0; i < att * per_attempt; ++i) rng_next(rs);
            s_ci[0] = (int)(rng_next(rs) % (uint32_t)U);
        }
        __syncthreads();
        // dist[i] = min(dist[i], |x_i - x_c|^2) (aux), with per-row and per-warp totals for the prefix search
        auto update_pass = [&](uint32_t ckey, bool first_centre) {
            unsigned long long wt = 0;
            for (int r0 = 0; r0 < rows; r0 += 4) {   // four rows per step: the key loads (L2) are all in flight first
                uint32_t kv[4], dv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = wbase + (r0 + u) * 32 + lane;
                    const bool ok = r0 + u < rows && i < U;
                    kv[u] = ok ? key_at(i) : ckey;               // distance 0 for the padding
                    dv[u] = (ok && !first_centre) ? aux[i] : (ok ? 0xffffffffu : 0u);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (r0 + u >= rows) break;   // warp-uniform
                    const int i = wbase + (r0 + u) * 32 + lane;
                    const uint32_t d = min(idist2(kv[u], ckey), dv[u]);
                    if (i < U) aux[i] = d;
                    const uint32_t rs32 = __reduce_add_sync(FULL, d);
                    if (IN_SMEM && lane == 0) s_rowsum[IN_SMEM ? warp : 0][IN_SMEM ? r0 + u : 0] = rs32;
                    wt += rs32;
                }
            }
            if (lane == 0) s_wtot[warp] = wt;
        };
        {
            const uint32_t ckey = key_at(s_ci[0]);
            if (tid == 0) {
                float r, g, b;
                unpackf(ckey, r, g, b);
                s_c[0] = make_float4(r, g, b, 0.f);
            }
            update_pass(ckey, true);
        }
        __syncthreads();
        unsigned long long sum0 = 0;
        for (int w = 0; w < FW; ++w) sum0 += s_wtot[w];
        for (int k = 1; k < K; ++k) {
            // the three trial positions of this step (the draws do not depend on the trials' outcome)
            if (tid == 0) {
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const uint32_t hi = rng_next(rs);
                    const unsigned long long v = ((unsigned long long)hi << 32) | rng_next(rs);
                    s_p[t] = __dmul_rn(__dmul_rn((double)v, 5.4210108624275221700372640043497e-20), (double)sum0);
                    s_ci[t] = U - 1;
                }
            }
            __syncthreads();
            // first i with inclusive prefix >= p  (== the sequential "p -= d; if (p <= 0) break")
            {
                unsigned long long before = 0;
                for (int w = 0; w < warp; ++w) before += s_wtot[w];
                const unsigned long long mytot = s_wtot[warp];
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const double p = s_p[t];
                    if (!((double)(before + mytot) >= p && ((double)before < p || warp == 0))) continue;  // warp-uniform
                    unsigned long long run = before;
                    int r = 0;
                    if (IN_SMEM) {
                        // row of the hit from the row totals (one scan per 32 rows), then one scan inside the row
                        for (int rb = 0; rb < rows; rb += 32) {
                            const uint32_t v = (rb + lane < rows) ? s_rowsum[IN_SMEM ? warp : 0][IN_SMEM ? rb + lane : 0] : 0u;
                            unsigned long long inc = v;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const unsigned long long nn = __shfl_up_sync(FULL, inc, o);
                                if (lane >= o) inc += nn;
                            }
                            const uint32_t hit = __ballot_sync(FULL, (double)(run + inc) >= p && rb + lane < rows);
                            if (hit) {
                                const int l = __ffs(hit) - 1;
                                r = rb + l;
                                run += __shfl_sync(FULL, inc, l) - __shfl_sync(FULL, (unsigned long long)v, l);
                                break;
                            }
                            run += __shfl_sync(FULL, inc, 31);
                            r = rb + 32;
                        }
                    }
                    for (; r < rows; ++r) {
                        const int i = wbase + r * 32 + lane;
                        const uint32_t v = i < U ? aux[i] : 0u;
                        uint32_t inc = v;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t nn = __shfl_up_sync(FULL, inc, o);
                            if (lane >= o) inc += nn;
                        }
                        const uint32_t rowsum = __shfl_sync(FULL, inc, 31);
                        if ((double)(run + rowsum) >= p) {
                            const uint32_t hit = __ballot_sync(FULL, (double)(run + inc) >= p);
                            const int idx = wbase + r * 32 + (__ffs(hit) - 1);
                            if (lane == 0 && idx < U - 1) atomicMin(&s_ci[t], idx);
                            break;
                        }
                        run += rowsum;
                    }
                }
            }
            __syncthreads();
            // one pass evaluates the three trials: s_t = sum_i min(|x_i - x_ci(t)|^2, dist[i])
            const int c0 = s_ci[0], c1 = s_ci[1], c2 = s_ci[2];
            const uint32_t t0 = key_at(c0), t1 = key_at(c1), t2 = key_at(c2);
            unsigned long long a0 = 0, a1 = 0, a2 = 0;
            for (int rb = 0; rb < rows; rb += 64) {   // 64 rows * 195075 < 2^32: 32-bit partial sums
                const int re = min(rows, rb + 64);
                uint32_t p0 = 0, p1 = 0, p2 = 0;
                for (int r0 = rb; r0 < re; r0 += 4) {   // four rows per step: loads first
                    uint32_t kv[4], dv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = wbase + (r0 + u) * 32 + lane;
                        const bool ok = r0 + u < re && i < U;
                        kv[u] = ok ? key_at(i) : 0u;
                        dv[u] = ok ? aux[i] : 0u;   // min(., 0) = 0 for the padding
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        p0 += min(idist2(kv[u], t0), dv[u]);
                        p1 += min(idist2(kv[u], t1), dv[u]);
                        p2 += min(idist2(kv[u], t2), dv[u]);
                    }
                }
                a0 += p0;
                a1 += p1;
                a2 += p2;
            }
            a0 = warp_sum_u64(a0);
            a1 = warp_sum_u64(a1);
            a2 = warp_sum_u64(a2);
            if (lane == 0) {
                s_red3[warp][0] = a0;
                s_red3[warp][1] = a1;
                s_red3[warp][2] = a2;
            }
            __syncthreads();
            unsigned long long best_s = ~0ull;
            int best_c = -1;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                unsigned long long s = 0;
                for (int w = 0; w < FW; ++w) s += s_red3[w][t];
                if (s < best_s) {   // strictly smaller: the first trial wins ties
                    best_s = s;
                    best_c = t == 0 ? c0 : t == 1 ? c1 : c2;
                }
            }
            const uint32_t bk = key_at(best_c);
            if (tid == 0) {
                float r, g, b;
                unpackf(bk, r, g, b);
                s_c[k] = make_float4(r, g, b, 0.f);
            }
            if (k + 1 < K) update_pass(bk, false);   // dist / totals for the next step (not needed after the last)
            sum0 = best_s;
            __syncthreads();
        }
        it = 1;
    }

    // ------------------------------ Lloyd ------------------------------------------
    // Assignment with Hamerly bounds.  Every point carries (in its aux word, free after the seeding)
    // an upper bound `ub` on its distance to its own centre and a lower bound `lb` on its distance to
    // every other centre, as 16-bit fixed point (1/64 colour units, ub rounded up, lb down, one extra
    // unit of slack each).  When the centres move by delta_k, ub += delta_own and lb -= max delta_other
    // stay valid; while ub < max(lb, half the distance from the own centre to its nearest neighbour)
    // the point provably keeps its label -- and by a margin (>= 1/64) far above the float32 rounding of
    // cv2's squared distances (< 3e-4 at 442), so cv2's argmin gives the same label.  Every other point
    // takes cv2's exact float32 search (separate mul/add, strict '<').  Sums are kept as exact integers
    // and updated incrementally when a label changes (== cv2's sequential float32 sums below 2^24).
    const long long t_seeded = clock64();
    // Ownership is ROW-INTERLEAVED here (row r of warp w = points (r * FW + w) * 32 ...): the list is
    // sorted by colour, so the points near a cluster boundary are neighbours in it; interleaving
    // spreads them over all warps (the seeding above needs contiguous blocks for its prefix search).
    for (bool first = true;; first = false) {
        if (!first && t_first == 0) t_first = clock64();
        if (tid < KMAX) s_asg[tid] = s_c[tid];
        if (first) {
            if (tid < KMAX * 4) (&s_sum[0][0])[tid] = 0;
            __syncthreads();
            uint32_t acc_lo[KC], acc_hi[KC];  // lo = G<<16 | B ; hi = count<<16 | R   (16-bit fields)
#pragma unroll
            for (int k = 0; k < KC; ++k) acc_lo[k] = acc_hi[k] = 0u;
            for (int r0 = 0; r0 < rows; r0 += 256) {
                const int r1 = min(rows, r0 + 256);
                for (int r4 = r0; r4 < r1; r4 += 4) {
                  // the key loads (L2) of four rows are in flight before the first search starts
                  uint32_t kv[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                      const int i = ((r4 + u) * FW + warp) * 32 + lane;
                      kv[u] = (r4 + u < r1 && i < U) ? key_at(i) : 0u;
                  }
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const int r = r4 + u;
                    int i = (r * FW + warp) * 32 + lane;
                    if (r >= r1 || i >= U) continue;
                    uint32_t key = kv[u];
                    float fr, fg, fb;
                    unpackf(key, fr, fg, fb);
                    float bd = fdist4(fr, fg, fb, s_c[0]), sd = 3e38f;
                    int bl = 0;
#pragma unroll
                    for (int k = 1; k < KC; ++k) {
                        if (k < K) {
                            float d = fdist4(fr, fg, fb, s_c[k]);
                            if (d < bd) {
                                sd = bd;
                                bd = d;
                                bl = k;
                            } else {
                                sd = fminf(sd, d);
                            }
                        }
                    }
                    aux[i] = F::pack((uint32_t)bl, F::q_up(bound_sqrt(bd)), F::q_dn(bound_sqrt(sd)));
                    const uint32_t plo = key & 0xffffu, pr = (key >> 16) | 0x10000u;
                    const uint32_t lo = ((plo & 0xff00u) << 8) | (plo & 0xffu);
#pragma unroll
                    for (int k = 0; k < KC; ++k) {
                        const bool m = (bl == k);
                        acc_lo[k] += m ? lo : 0u;
                        acc_hi[k] += m ? pr : 0u;
                    }
                  }
                }
                // flush the 16-bit fields before they can overflow (256 rows * 255 < 65536)
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    if (k < K) {
                        uint32_t sb = __reduce_add_sync(FULL, acc_lo[k] & 0xffffu);
                        uint32_t sg = __reduce_add_sync(FULL, acc_lo[k] >> 16);
                        uint32_t sr = __reduce_add_sync(FULL, acc_hi[k] & 0xffffu);
                        uint32_t sn = __reduce_add_sync(FULL, acc_hi[k] >> 16);
                        if (lane == 0 && sn) {
                            atomicAdd(&s_sum[k][0], (int)sr);
                            atomicAdd(&s_sum[k][1], (int)sg);
                            atomicAdd(&s_sum[k][2], (int)sb);
                            atomicAdd(&s_sum[k][3], (int)sn);
                        }
                    }
                    acc_lo[k] = acc_hi[k] = 0u;
                }
            }
        } else {
            __syncthreads();  // s_tab of the last update is visible
            uint8_t* q = s_queue[warp];
            // rows of this warp that are completely inside the list need no bounds check
            const int full_rows_all = U >> 5;
            const int r_full = full_rows_all > warp ? (full_rows_all - warp + FW - 1) / FW : 0;
            // Two phases per block of QROWS rows so that the expensive path runs on full warps: (A) every
            // lane tests the bounds of its points and the failing ones are compacted into the warp's
            // queue; (B) the queue is drained 32 points at a time.
            Q2* q2 = s_queue2[warp];
            int q2n = 0, q2head = 0;
            // (B2) cv2's exact search for one point (rl = row of this warp << 5 | lane)
            auto search_point = [&](int rl) {
                const int i = ((rl >> 5) * FW + warp) * 32 + (rl & 31);
                const uint32_t a = F::label(aux[i]);
                const uint32_t key = key_at(i);
                float fr, fg, fb;
                unpackf(key, fr, fg, fb);
                if (P.dbg) atomicAdd(&s_cnt[1], 1u);
                float bd = fdist4(fr, fg, fb, s_c[0]), sd = 3e38f;
                int bl = 0;
#pragma unroll
                for (int k = 1; k < KC; ++k) {
                    if (k < K) {
                        float d = fdist4(fr, fg, fb, s_c[k]);
                        if (d < bd) {
                            sd = bd;
                            bd = d;
                            bl = k;
                        } else {
                            sd = fminf(sd, d);
                        }
                    }
                }
                aux[i] = F::pack((uint32_t)bl, F::q_up(bound_sqrt(bd)), F::q_dn(bound_sqrt(sd)));
                if (P.dbg && (uint32_t)bl != a) atomicAdd(&s_cnt[2], 1u);
                if ((uint32_t)bl != a) {
                    const int cr = (int)(key >> 16), cg = (int)((key >> 8) & 255u), cb = (int)(key & 255u);
                    atomicAdd(&s_sum[a][0], -cr);
                    atomicAdd(&s_sum[a][1], -cg);
                    atomicAdd(&s_sum[a][2], -cb);
                    atomicAdd(&s_sum[a][3], -1);
                    atomicAdd(&s_sum[bl][0], cr);
                    atomicAdd(&s_sum[bl][1], cg);
                    atomicAdd(&s_sum[bl][2], cb);
                    atomicAdd(&s_sum[bl][3], 1);
                }
            };
            for (int rb0 = 0; rb0 < rows; rb0 += QROWS) {
                const int rb1 = min(rows, rb0 + QROWS);
                int qn = 0;
                if (rb0 + QROWS <= r_full) {
                    // the whole block lies inside the list (warp-uniform): the eight state words are loaded first, the
                    // row addresses are compile-time offsets from one pointer, no validity tests
                    uint32_t* const ap = aux + (rb0 * FW + warp) * 32 + lane;
                    const uint32_t below = (1u << lane) - 1u;
                    uint32_t xs[QROWS];
#pragma unroll
                    for (int u = 0; u < QROWS; ++u) xs[u] = ap[u * FW * 32];
#pragma unroll
                    for (int u = 0; u < QROWS; ++u) {
                        const uint32_t x = xs[u];
                        const uint32_t a = F::label(x);
                        const uint4 t = s_tab[a];
                        const uint32_t ub = min(F::UMAX, F::ub(x) + t.x);
                        const uint32_t lb = max(F::lb(x), t.y) - t.y;
                        ap[u * FW * 32] = F::pack(a, ub, lb);
                        const bool need = ub >= max(lb, t.z);
                        const uint32_t bal = __ballot_sync(FULL, need);
                        if (need) q[qn + __popc(bal & below)] = (uint8_t)(u * 32 + lane);
                        qn += __popc(bal);
                    }
                } else
                for (int r = rb0; r < rb1; ++r) {
                    const int i = (r * FW + warp) * 32 + lane;
                    bool need = false;
                    if (r < r_full || i < U) {   // first test is warp-uniform and almost always true
                        const uint32_t x = aux[i];
                        const uint32_t a = F::label(x);
                        const uint4 t = s_tab[a];
                        const uint32_t ub = min(F::UMAX, F::ub(x) + t.x);
                        const uint32_t lb = max(F::lb(x), t.y) - t.y;
                        aux[i] = F::pack(a, ub, lb);
                        need = ub >= max(lb, t.z);
                    }
                    const uint32_t bal = __ballot_sync(FULL, need);
                    if (need) q[qn + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)((r - rb0) * 32 + lane);
                    qn += __popc(bal);
                }
                if (P.dbg && lane == 0) atomicAdd(&s_cnt[0], (unsigned)qn);
                __syncwarp();
                // (B1) tighten the upper bound of the queued points with one exact distance to their own centre; the
                // ones that still fail go to a second queue that is shared by all blocks of this warp's rows, so that
                // the full search (B2) runs on whole warps: per block it would find ~6 of 32 lanes busy
                for (int j0 = 0; j0 < qn; j0 += 32) {   // warp-uniform trip count
                    const int j = j0 + lane;
                    bool full = false;
                    int rl = 0;
                    if (j < qn) {
                        const int qe = (int)q[j];
                        rl = ((rb0 + (qe >> 5)) << 5) | (qe & 31);          // row of this warp, lane
                        const int i = ((rb0 + (qe >> 5)) * FW + warp) * 32 + (qe & 31);
                        const uint32_t x = aux[i];
                        const uint32_t a = F::label(x), lb = F::lb(x);
                        const uint32_t bound = max(lb, s_tab[a].z);
                        float fr, fg, fb;
                        unpackf(key_at(i), fr, fg, fb);
                        const uint32_t ub = F::q_up(bound_sqrt(fdist4(fr, fg, fb, s_c[a])));
                        if (ub < bound) aux[i] = F::pack(a, ub, lb);
                        else full = true;
                    }
                    const uint32_t bal = __ballot_sync(FULL, full);
                    if (full) q2[(q2head + q2n + __popc(bal & ((1u << lane) - 1u))) & 63] = (Q2)rl;
                    q2n += __popc(bal);
                    __syncwarp();
                    if (q2n >= 32) {
                        search_point((int)q2[(q2head + lane) & 63]);
                        q2head = (q2head + 32) & 63;
                        q2n -= 32;
                        __syncwarp();
                    }
                }
                __syncwarp();
            }
            if (lane < q2n) search_point((int)q2[(q2head + lane) & 63]);
        }
        __syncthreads();
        // cv2's float32 centre sums: the exact integers while every channel sum is below 2^24 (always, for lists that
        // fit in shared memory: 57 344 x 255 < 2^24), else the sequential float32 sum reproduced by seq_f32_sums3
        if (tid < K * 3) s_fsum[tid / 3][tid % 3] = (float)s_sum[tid / 3][tid % 3];
        if (!IN_SMEM) {
            bool any_long = false;
            for (int k = 0; k < K; ++k)
                any_long |= s_sum[k][0] >= (1 << 24) || s_sum[k][1] >= (1 << 24) || s_sum[k][2] >= (1 << 24);
            if (any_long) {   // block-uniform (shared memory)
                if (tid == 0) s_flag = 1;
                __syncthreads();
                for (int k = 0; k < K; ++k)
                    if (s_sum[k][0] >= (1 << 24) || s_sum[k][1] >= (1 << 24) || s_sum[k][2] >= (1 << 24))
                        seq_f32_sums3<F, FT>(k, U, aux, keys, *reinterpret_cast<SeqShared*>(&s_seq), s_fsum[k]);
            }
        }
        __syncthreads();
        // empty-cluster repair (cv2: the biggest cluster gives up its farthest member, last max wins)
        for (int k = 0; k < K; ++k) {
            if (s_sum[k][3] != 0) continue;  // uniform (shared memory)
            int mk = 0;
            for (int k1 = 1; k1 < K; ++k1)
                if (s_sum[mk][3] < s_sum[k1][3]) mk = k1;
            const float sc = __fdiv_rn(1.f, (float)s_sum[mk][3]);
            const float4 base = make_float4(__fmul_rn(s_fsum[mk][0], sc), __fmul_rn(s_fsum[mk][1], sc),
                                            __fmul_rn(s_fsum[mk][2], sc), 0.f);
            if (tid == 0) s_far = 0ull;
            __syncthreads();
            unsigned long long best = 0ull;
            for (int r = 0; r < rows; ++r) {
                int i = (r * FW + warp) * 32 + lane;
                if (i >= U) continue;
                if ((int)F::label(aux[i]) != mk) continue;
                float fr, fg, fb;
                unpackf(key_at(i), fr, fg, fb);
                float d = fdist4(fr, fg, fb, base);
                unsigned long long cand = (((unsigned long long)__float_as_uint(d) << 32) | (uint32_t)i) + 1ull;
                best = cand > best ? cand : best;
            }
            if (best) atomicMax(&s_far, best);
            __syncthreads();
            if (tid == 0) {
                int far = (int)(uint32_t)((s_far - 1ull) & 0xffffffffull);
                uint32_t fk = key_at(far);
                aux[far] = F::pack((uint32_t)k, F::UMAX, 0u);  // bounds unknown: ub = inf, lb = 0
                s_c[mk] = base;  // OpenCV stores the donor's provisional mean in old_centers[max_k]
                s_sum[mk][0] -= (int)(fk >> 16);
                s_sum[mk][1] -= (int)((fk >> 8) & 255u);
                s_sum[mk][2] -= (int)(fk & 255u);
                s_sum[mk][3] -= 1;
                s_sum[k][0] += (int)(fk >> 16);
                s_sum[k][1] += (int)((fk >> 8) & 255u);
                s_sum[k][2] += (int)(fk & 255u);
                s_sum[k][3] += 1;
                // cv2 moves the point in float32: center[max_k] -= sample, center[k] += sample
                const float xr = (float)(fk >> 16), xg = (float)((fk >> 8) & 255u), xb = (float)(fk & 255u);
                s_fsum[mk][0] = __fsub_rn(s_fsum[mk][0], xr);
                s_fsum[mk][1] = __fsub_rn(s_fsum[mk][1], xg);
                s_fsum[mk][2] = __fsub_rn(s_fsum[mk][2], xb);
                s_fsum[k][0] = __fadd_rn(s_fsum[k][0], xr);
                s_fsum[k][1] = __fadd_rn(s_fsum[k][1], xg);
                s_fsum[k][2] = __fadd_rn(s_fsum[k][2], xb);
            }
            __syncthreads();
        }
        // new centres
        if (tid < K) {
            s_old[tid] = s_c[tid];
            const float sc = __fdiv_rn(1.f, (float)s_sum[tid][3]);
            s_c[tid] = make_float4(__fmul_rn(s_fsum[tid][0], sc), __fmul_rn(s_fsum[tid][1], sc),
                                   __fmul_rn(s_fsum[tid][2], sc), 0.f);
        }
        __syncthreads();
        double shift = 0.0;
        const bool first_seeded = (P.init != nullptr && it == 0);
        for (int k = 0; k < K; ++k) {
            const float4 c = s_c[k], o = s_old[k];
            double t0 = (double)__fsub_rn(c.x, o.x), t1 = (double)__fsub_rn(c.y, o.y), t2 = (double)__fsub_rn(c.z, o.z);
            double s = __dmul_rn(t0, t0);
            s = __dadd_rn(s, __dmul_rn(t1, t1));
            s = __dadd_rn(s, __dmul_rn(t2, t2));
            shift = fmax(shift, s);
        }
        ++it;
        const int last_it = P.max_iter > 2 ? P.max_iter : 2;
        const bool last = (it == last_it) || (!first_seeded && shift <= P.eps2);
        if (last) break;
        // bounds bookkeeping for the next assignment: how far every centre moved since the assignment
        // just done (s_asg), and half the distance of every centre to its nearest neighbour
        uint32_t hq_mine = 0u;
        if (tid < K) {
            const float4 c = s_c[tid], o = s_asg[tid];
            const float dx = c.x - o.x, dy = c.y - o.y, dz = c.z - o.z;
            s_dq[tid] = F::q_up(sqrtf(dx * dx + dy * dy + dz * dz));
            float nn = 3e38f;
            for (int j = 0; j < K; ++j) {
                if (j == tid) continue;
                const float4 e = s_c[j];
                const float ex = c.x - e.x, ey = c.y - e.y, ez = c.z - e.z;
                nn = fminf(nn, ex * ex + ey * ey + ez * ez);
            }
            hq_mine = F::q_dn(0.5f * sqrtf(nn));
        }
        __syncthreads();
        if (tid < K) {
            // largest and second largest drift (every one of the K threads scans the K values itself)
            uint32_t m1 = 0, m2 = 0, km = 0;
            for (int k = 0; k < K; ++k) {
                const uint32_t d = s_dq[k];
                if (d > m1) {
                    m2 = m1;
                    m1 = d;
                    km = k;
                } else if (d > m2) {
                    m2 = d;
                }
            }
            s_tab[tid] = make_uint4(s_dq[tid], (uint32_t)tid == km ? m2 : m1, hq_mine, 0u);
        }
    }
    const long long t_conv = clock64();
    // compactness with the final centres and the labels of the last assignment; labels out
    uint8_t* labels = P.labels + slot * P.max_unique;
    double part = 0.0;
    for (int r4 = 0; r4 < rows; r4 += 4) {
        uint32_t kv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = ((r4 + u) * FW + warp) * 32 + lane;
            kv[u] = (r4 + u < rows && i < U) ? key_at(i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {   // same order of additions as a row-by-row walk
            const int i = ((r4 + u) * FW + warp) * 32 + lane;
            if (r4 + u >= rows || i >= U) continue;
            const uint32_t a = F::label(aux[i]);
            float fr, fg, fb;
            unpackf(kv[u], fr, fg, fb);
            part += (double)fdist4(fr, fg, fb, s_c[a]);
            labels[i] = (uint8_t)a;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
    if (lane == 0) s_redd[warp] = part;
    __syncthreads();
    if (tid == 0) {
        double comp = 0.0;
        for (int w = 0; w < FW; ++w) comp += s_redd[w];
        P.compact[slot] = comp;
        P.iters[slot] = it;
        P.inexact[slot] = s_flag;
        if (P.dbg) {
            unsigned long long* d = P.dbg + slot * 8;
            const long long t_end = clock64();
            if (t_first == 0) t_first = t_conv;
            d[0] = t_loaded - t_start;
            d[1] = t_seeded - t_loaded;
            d[2] = t_first - t_seeded;
            d[3] = t_conv - t_first;
            d[4] = t_end - t_conv;
            d[5] = it;
            d[6] = U;
            d[7] = t_end - t_start;
            d[0] = ((unsigned long long)s_cnt[0] << 40) | ((unsigned long long)s_cnt[1] << 20) | s_cnt[2];
        }
    }
    if (tid < K) {
        const float4 c = s_c[tid];
        float* o = P.centers + slot * KMAX * 3 + tid * 3;
        o[0] = c.x;
        o[1] = c.y;
        o[2] = c.z;
        for (int j = 0; j < 4; ++j) P.sums[slot * KMAX * 4 + tid * 4 + j] = (unsigned long long)s_sum[tid][j];
    }
}

// order[r] = the image with the r-th longest colour list (ties by index): the block scheduler hands out CTAs in
// grid order, so the long lists start first and the short ones fill the tail of the launch
__global__ void __launch_bounds__(256) k_kmeans_order(const int32_t* __restrict__ count, int n, int32_t* __restrict__ order) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int ci = count[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
        const int cj = count[j];
        rank += (cj > ci) || (cj == ci && j < i);
    }
    order[rank] = i;
}

template <int KC>
int launch_kc(llfe_ctx* ctx, const KmParams& P, int n, int pts2, int pts1) {
    // the limits must not depend on this call's max_unique (pts1 is capped by it): a later call may bring longer lists
    if (llfe_first_use(ctx, (const void*)k_kmeans_fast<KC, true, FT_STD>)) {
        LLFE_CUDA(cudaFuncSetAttribute(k_kmeans_fast<KC, true, FT_STD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(ctx->smem_optin - STATIC_SMEM_BOUND)));
        LLFE_CUDA(cudaFuncSetAttribute(k_kmeans_fast<KC, true, FT_LONG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(ctx->smem_optin - STATIC_SMEM_BOUND_LONG)));
    }
    // One launch for the whole batch: the block scheduler back-fills SMs as CTAs finish, so attempts
    // that need many iterations do not hold up a wave.  Lists of up to pts2 colours run two CTAs per SM.
    if (P.order) {
        LLFE_KERNEL(ctx, "k_kmeans_order");
        k_kmeans_order<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(P.count, n, const_cast<int32_t*>(P.order));
        LLFE_LAUNCHED(ctx);
    }
    LLFE_KERNEL(ctx, "k_kmeans_fast");
    k_kmeans_fast<KC, true, FT_STD><<<dim3(P.attempts, n), FT_STD, (size_t)pts2 * 4, ctx->stream>>>(P, 0, pts2, 0);
    LLFE_LAUNCHED(ctx);
    // longer lists: one CTA of 1024 threads per SM with all of its shared memory (every CTA of the other images exits
    // at once)
    if (P.max_unique > pts2 && pts1 > pts2) {
        LLFE_KERNEL(ctx, "k_kmeans_fast_long");
        k_kmeans_fast<KC, true, FT_LONG><<<dim3(P.attempts, n), FT_LONG, (size_t)pts1 * 4, ctx->stream>>>(P, pts2, pts1, 0);
        LLFE_LAUNCHED(ctx);
    }
    // lists that do not fit in shared memory at all: global scratch, P.dist_images images per launch
    if (P.max_unique > pts1) {
        for (int i0 = 0; i0 < n; i0 += P.dist_images) {
            const int m = (n - i0) < P.dist_images ? (n - i0) : P.dist_images;
            LLFE_KERNEL(ctx, "k_kmeans_fast_global");
            k_kmeans_fast<KC, false, FT_LONG><<<dim3(P.attempts, m), FT_LONG, 0, ctx->stream>>>(P, 0, pts1, i0);   // one CTA per SM: 32 warps on the list
            LLFE_LAUNCHED(ctx);
        }
    }
    return LLFE_OK;
}

}  // namespace

// pts2 / pts1: how many colours (4 bytes each) fit in the dynamic shared memory of one CTA with two / one CTA per SM
int launch_kmeans_fast(llfe_ctx* ctx, const KmParams& P0, int n) {
    KmParams P = P0;
    // phase clocks [n][attempts][8] u64, only into a buffer registered (and validated) by llfe_set_debug_buffer
    P.dbg = (ctx->dbg_kmeans && ctx->dbg_kmeans_bytes >= (size_t)n * P.attempts * 8 * 8) ? ctx->dbg_kmeans : nullptr;
    const size_t per_cta2 = (ctx->smem_optin + 1024) / 2 - 1024;   // two CTAs per SM, 1 KB reserved per CTA
    auto points = [&](size_t per_cta, size_t static_smem) {
        size_t avail = per_cta > static_smem ? per_cta - static_smem : 0;
        long long p = (long long)(avail / 4);
        if (p > P.max_unique) p = P.max_unique;
        if (p > RMAX_POINTS) p = RMAX_POINTS;
        return (int)(p & ~31ll);
    };
    const int pts2 = points(per_cta2, STATIC_SMEM_BOUND), pts1 = points(ctx->smem_optin, STATIC_SMEM_BOUND_LONG);
    if (P.k <= 5) return launch_kc<5>(ctx, P, n, pts2, pts1);
    if (P.k <= 8) return launch_kc<8>(ctx, P, n, pts2, pts1);
    if (P.k <= 16) return launch_kc<16>(ctx, P, n, pts2, pts1);
    return launch_kc<32>(ctx, P, n, pts2, pts1);
}
