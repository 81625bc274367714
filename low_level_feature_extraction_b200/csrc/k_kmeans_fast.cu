// Fast path of cv2.kmeans over UNWEIGHTED colour lists (the reference's palette call,
// color_extractor.py:189-196): same arithmetic and results as k_kmeans (k_kmeans.cu), laid
// out for the SM:
//   * one CTA per (attempt, image); the colour list lives in SHARED MEMORY for the whole
//     solve (key | label<<24 per point, plus one aux word per point: the kmeans++ distance),
//     so no iteration touches HBM or L2;
//   * warp-blocked point ownership (warp w owns a contiguous block, lanes stride it): bank-
//     conflict-free and contiguous for the kmeans++ prefix search;
//   * centroid partial sums in packed per-thread REGISTER accumulators
//     (count|R and G|B as 16-bit fields), combined once per iteration with REDUX
//     warp reductions -- no atomics in the inner loop;
//   * lists too long for shared memory fall back to a global scratch copy (same code).
#include "llfe_common.cuh"
#include "llfe_device.cuh"
#include "k_kmeans_shared.cuh"

namespace {

constexpr int FT = 512;
constexpr int FW = FT / 32;

__device__ __forceinline__ float4 ld_center(const float4* c, int k) { return c[k]; }

__device__ __forceinline__ float fdist4(float r, float g, float b, float4 c) {
    float t0 = __fsub_rn(r, c.x), t1 = __fsub_rn(g, c.y), t2 = __fsub_rn(b, c.z);
    float d = __fmul_rn(t0, t0);
    d = __fadd_rn(d, __fmul_rn(t1, t1));
    d = __fadd_rn(d, __fmul_rn(t2, t2));
    return d;
}

__device__ __forceinline__ void unpackf(uint32_t key, float& r, float& g, float& b) {
    r = (float)((key >> 16) & 255u);
    g = (float)((key >> 8) & 255u);
    b = (float)(key & 255u);
}

template <int KC>
__global__ void __launch_bounds__(FT, 1) k_kmeans_fast(KmParams P, int smem_points) {
    extern __shared__ uint32_t dyn[];
    const int att = blockIdx.x, img = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int U = min(P.count[img], P.max_unique);
    const int K = min(P.k, U);
    const size_t slot = (size_t)img * P.attempts + att;
    if (K <= 1) {
        if (tid == 0) {
            P.compact[slot] = 0.0;
            P.iters[slot] = 0;
            P.inexact[slot] = 0;
        }
        return;
    }
    const uint32_t* keys = P.keys + (size_t)img * P.max_unique;
    // point storage: shared memory when the list fits, else this slot's global scratch
    uint32_t* pts;
    uint32_t* aux;
    if (U <= smem_points) {
        pts = dyn;
        aux = dyn + smem_points;
    } else {
        pts = P.dist + slot * 2 * (size_t)P.max_unique;
        aux = pts + P.max_unique;
    }
    // warp-blocked ownership: rows of 32 consecutive points, `rows` rows per warp
    const int rows = (U + 32 * FW - 1) / (32 * FW);
    const int wbase = warp * rows * 32;

    __shared__ float4 s_c[KMAX];
    __shared__ float4 s_old[KMAX];
    __shared__ unsigned long long s_tot[KMAX][4];
    __shared__ unsigned long long s_red[FW];
    __shared__ unsigned long long s_wtot[FW];
    __shared__ double s_redd[FW];
    __shared__ int s_ci;
    __shared__ double s_p;
    __shared__ int s_flag;
    __shared__ unsigned long long s_far;

    if (tid == 0) s_flag = 0;
    for (int r = 0; r < rows; ++r) {
        int i = wbase + r * 32 + lane;
        if (i < U) pts[i] = keys[i] & 0xffffffu;
    }
    __syncthreads();

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
            s_ci = (int)(rng_next(rs) % (uint32_t)U);
        }
        __syncthreads();
        uint32_t ckey = pts[s_ci];
        if (tid == 0) {
            float r, g, b;
            unpackf(ckey, r, g, b);
            s_c[0] = make_float4(r, g, b, 0.f);
        }
        unsigned long long part = 0;
        for (int r = 0; r < rows; ++r) {
            int i = wbase + r * 32 + lane;
            if (i < U) {
                uint32_t d = idist(pts[i], ckey);
                aux[i] = d;
                part += d;
            }
        }
        part = warp_sum_u64(part);
        __syncthreads();  // everyone has read s_ci
        if (lane == 0) s_wtot[warp] = part;
        __syncthreads();
        unsigned long long sum0 = 0;
        for (int w = 0; w < FW; ++w) sum0 += s_wtot[w];
        for (int k = 1; k < K; ++k) {
            unsigned long long best_s = ~0ull;
            int best_c = -1;
            unsigned long long before = 0;
            for (int w = 0; w < warp; ++w) before += s_wtot[w];
            const unsigned long long mytot = s_wtot[warp];
            for (int trial = 0; trial < 3; ++trial) {
                if (tid == 0) {
                    uint32_t t = rng_next(rs);
                    unsigned long long v = ((unsigned long long)t << 32) | rng_next(rs);
                    s_p = __dmul_rn(__dmul_rn((double)v, 5.4210108624275221700372640043497e-20), (double)sum0);
                    s_ci = U - 1;
                }
                __syncthreads();
                const double p = s_p;
                // first i with inclusive prefix >= p  (== the sequential "p -= d; if (p <= 0) break")
                if ((double)(before + mytot) >= p && ((double)before < p || warp == 0)) {
                    unsigned long long run = before;
                    for (int r = 0; r < rows; ++r) {
                        int i = wbase + r * 32 + lane;
                        uint32_t v = i < U ? aux[i] : 0u;
                        uint32_t inc = v;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            uint32_t nn = __shfl_up_sync(0xffffffffu, inc, o);
                            if (lane >= o) inc += nn;
                        }
                        uint32_t rowsum = __shfl_sync(0xffffffffu, inc, 31);
                        if ((double)(run + rowsum) >= p) {
                            uint32_t hit = __ballot_sync(0xffffffffu, (double)(run + inc) >= p);
                            int idx = wbase + r * 32 + (__ffs(hit) - 1);
                            if (lane == 0 && idx < U - 1) atomicMin(&s_ci, idx);
                            break;
                        }
                        run += rowsum;
                    }
                }
                __syncthreads();
                const int ci = s_ci;
                const uint32_t tk = pts[ci];
                unsigned long long ps = 0;
                for (int r = 0; r < rows; ++r) {
                    int i = wbase + r * 32 + lane;
                    if (i < U) ps += min(idist(pts[i], tk), aux[i]);
                }
                ps = warp_sum_u64(ps);
                if (lane == 0) s_red[warp] = ps;
                __syncthreads();
                unsigned long long s = 0;
                for (int w = 0; w < FW; ++w) s += s_red[w];
                if (s < best_s) {
                    best_s = s;
                    best_c = ci;
                }
                __syncthreads();
            }
            const uint32_t bk = pts[best_c];
            if (tid == 0) {
                float r, g, b;
                unpackf(bk, r, g, b);
                s_c[k] = make_float4(r, g, b, 0.f);
            }
            unsigned long long wt = 0;
            for (int r = 0; r < rows; ++r) {
                int i = wbase + r * 32 + lane;
                if (i < U) {
                    uint32_t d = min(idist(pts[i], bk), aux[i]);
                    aux[i] = d;
                    wt += d;
                }
            }
            wt = warp_sum_u64(wt);
            if (lane == 0) s_wtot[warp] = wt;
            sum0 = best_s;
            __syncthreads();
        }
        it = 1;
    }

    // ------------------------------ Lloyd ------------------------------------------
    for (;;) {
        if (tid < KMAX * 4) (&s_tot[0][0])[tid] = 0ull;
        __syncthreads();
        uint32_t acc_lo[KC], acc_hi[KC];  // lo = G<<16 | B ; hi = count<<16 | R   (16-bit fields)
#pragma unroll
        for (int k = 0; k < KC; ++k) acc_lo[k] = acc_hi[k] = 0u;
        for (int r0 = 0; r0 < rows; r0 += 256) {
            const int r1 = min(rows, r0 + 256);
            for (int r = r0; r < r1; ++r) {
                int i = wbase + r * 32 + lane;
                if (i >= U) continue;
                uint32_t key = pts[i] & 0xffffffu;
                float fr, fg, fb;
                unpackf(key, fr, fg, fb);
                float bd = fdist4(fr, fg, fb, s_c[0]);
                int bl = 0;
#pragma unroll
                for (int k = 1; k < KC; ++k) {
                    if (k < K) {
                        float d = fdist4(fr, fg, fb, s_c[k]);
                        if (d < bd) {
                            bd = d;
                            bl = k;
                        }
                    }
                }
                pts[i] = key | ((uint32_t)bl << 24);
                const uint32_t plo = key & 0xffffu, pr = (key >> 16) | 0x10000u;
                const uint32_t lo = ((plo & 0xff00u) << 8) | (plo & 0xffu);
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const bool m = (bl == k);
                    acc_lo[k] += m ? lo : 0u;
                    acc_hi[k] += m ? pr : 0u;
                }
            }
            // flush the 16-bit fields before they can overflow (256 rows * 255 < 65536)
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                if (k < K) {
                    uint32_t sb = __reduce_add_sync(0xffffffffu, acc_lo[k] & 0xffffu);
                    uint32_t sg = __reduce_add_sync(0xffffffffu, acc_lo[k] >> 16);
                    uint32_t sr = __reduce_add_sync(0xffffffffu, acc_hi[k] & 0xffffu);
                    uint32_t sn = __reduce_add_sync(0xffffffffu, acc_hi[k] >> 16);
                    if (lane == 0 && sn) {
                        atomicAdd(&s_tot[k][0], (unsigned long long)sr);
                        atomicAdd(&s_tot[k][1], (unsigned long long)sg);
                        atomicAdd(&s_tot[k][2], (unsigned long long)sb);
                        atomicAdd(&s_tot[k][3], (unsigned long long)sn);
                    }
                }
                acc_lo[k] = acc_hi[k] = 0u;
            }
        }
        __syncthreads();
        // empty-cluster repair (cv2: the biggest cluster gives up its farthest member, last max wins)
        for (int k = 0; k < K; ++k) {
            if (s_tot[k][3] != 0) continue;  // uniform (shared memory)
            int mk = 0;
            for (int k1 = 1; k1 < K; ++k1)
                if (s_tot[mk][3] < s_tot[k1][3]) mk = k1;
            const float sc = __fdiv_rn(1.f, (float)s_tot[mk][3]);
            const float4 base = make_float4(__fmul_rn((float)s_tot[mk][0], sc), __fmul_rn((float)s_tot[mk][1], sc),
                                            __fmul_rn((float)s_tot[mk][2], sc), 0.f);
            if (tid == 0) s_far = 0ull;
            __syncthreads();
            unsigned long long best = 0ull;
            for (int r = 0; r < rows; ++r) {
                int i = wbase + r * 32 + lane;
                if (i >= U) continue;
                uint32_t w = pts[i];
                if ((int)(w >> 24) != mk) continue;
                float fr, fg, fb;
                unpackf(w, fr, fg, fb);
                float d = fdist4(fr, fg, fb, base);
                unsigned long long cand = (((unsigned long long)__float_as_uint(d) << 32) | (uint32_t)i) + 1ull;
                best = cand > best ? cand : best;
            }
            if (best) atomicMax(&s_far, best);
            __syncthreads();
            if (tid == 0) {
                int far = (int)(uint32_t)((s_far - 1ull) & 0xffffffffull);
                uint32_t fk = pts[far] & 0xffffffu;
                pts[far] = fk | ((uint32_t)k << 24);
                s_c[mk] = base;  // OpenCV stores the donor's provisional mean in old_centers[max_k]
                s_tot[mk][0] -= fk >> 16;
                s_tot[mk][1] -= (fk >> 8) & 255u;
                s_tot[mk][2] -= fk & 255u;
                s_tot[mk][3] -= 1;
                s_tot[k][0] += fk >> 16;
                s_tot[k][1] += (fk >> 8) & 255u;
                s_tot[k][2] += fk & 255u;
                s_tot[k][3] += 1;
            }
            __syncthreads();
        }
        // new centres
        if (tid < K) {
            s_old[tid] = s_c[tid];
            const float sc = __fdiv_rn(1.f, (float)s_tot[tid][3]);
            if (s_tot[tid][0] >= (1ull << 24) || s_tot[tid][1] >= (1ull << 24) || s_tot[tid][2] >= (1ull << 24))
                atomicOr(&s_flag, 1);
            s_c[tid] = make_float4(__fmul_rn((float)s_tot[tid][0], sc), __fmul_rn((float)s_tot[tid][1], sc),
                                   __fmul_rn((float)s_tot[tid][2], sc), 0.f);
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
        __syncthreads();
    }
    // compactness with the final centres and the labels of the last assignment; labels out
    uint8_t* labels = P.labels + slot * P.max_unique;
    double part = 0.0;
    for (int r = 0; r < rows; ++r) {
        int i = wbase + r * 32 + lane;
        if (i >= U) continue;
        uint32_t w = pts[i];
        float fr, fg, fb;
        unpackf(w, fr, fg, fb);
        part += (double)fdist4(fr, fg, fb, s_c[w >> 24]);
        labels[i] = (uint8_t)(w >> 24);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_redd[warp] = part;
    __syncthreads();
    if (tid == 0) {
        double comp = 0.0;
        for (int w = 0; w < FW; ++w) comp += s_redd[w];
        P.compact[slot] = comp;
        P.iters[slot] = it;
        P.inexact[slot] = s_flag;
    }
    if (tid < K) {
        const float4 c = s_c[tid];
        float* o = P.centers + slot * KMAX * 3 + tid * 3;
        o[0] = c.x;
        o[1] = c.y;
        o[2] = c.z;
        for (int j = 0; j < 4; ++j) P.sums[slot * KMAX * 4 + tid * 4 + j] = s_tot[tid][j];
    }
}

}  // namespace

// smem_points: how many points (8 bytes each) the dynamic shared memory of one CTA can hold
int launch_kmeans_fast(llfe_ctx* ctx, const KmParams& P, int n) {
    static bool attr_set = false;
    const size_t static_smem = 4096;  // centres, totals, reduction scratch (upper bound)
    size_t avail = ctx->smem_optin > static_smem ? ctx->smem_optin - static_smem : 0;
    int smem_points = (int)(avail / 8);
    if (smem_points > P.max_unique) smem_points = P.max_unique;
    smem_points &= ~31;
    const size_t dyn = (size_t)smem_points * 8;
    if (!attr_set) {
        LLFE_CUDA(cudaFuncSetAttribute(k_kmeans_fast<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)avail));
        LLFE_CUDA(cudaFuncSetAttribute(k_kmeans_fast<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)avail));
        LLFE_CUDA(cudaFuncSetAttribute(k_kmeans_fast<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)avail));
        attr_set = true;
    }
    dim3 grid(P.attempts, n);
    LLFE_KERNEL(ctx, "k_kmeans_fast");
    if (P.k <= 8)
        k_kmeans_fast<8><<<grid, FT, dyn, ctx->stream>>>(P, smem_points);
    else if (P.k <= 16)
        k_kmeans_fast<16><<<grid, FT, dyn, ctx->stream>>>(P, smem_points);
    else
        k_kmeans_fast<32><<<grid, FT, dyn, ctx->stream>>>(P, smem_points);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}
