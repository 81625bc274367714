// k-means over colour lists (cv2.kmeans as called at color_extractor.py:189-196).
//
// One CTA per (attempt, image): kmeans++ seeding with cv::RNG (3 trials per
// centre), Lloyd iterations to convergence, compactness; a second tiny kernel
// keeps the best attempt.  Nearest-centre assignment uses OpenCV's float32
// distance (separate multiplies and adds, strict '<').  Centroid sums are exact
// integers (warp-private shared-memory accumulators -> block reduction), which
// equals cv2's sequential float32 sums bit for bit while every cluster's channel
// sum stays below 2^24 (always true for unique-colour lists of design images;
// `inexact` reports when it is not).
//
// The same kernel runs the "seeded" mode (given initial centres, optional
// per-colour weights = pixel counts, optional exact-sum centre rule
// c = float(double(sum)/double(count))) that backs the per-pixel k-means mode.
#include "llfe_common.cuh"
#include "llfe_device.cuh"
#include "k_kmeans_shared.cuh"

namespace {

constexpr int KT = 512;  // threads per CTA
constexpr int KW = KT / 32;

__device__ __forceinline__ void unpack(uint32_t key, float& r, float& g, float& b) {
    r = (float)(key >> 16);
    g = (float)((key >> 8) & 255u);
    b = (float)(key & 255u);
}

// OpenCV normL2Sqr for 3 floats: ((0 + t0*t0) + t1*t1) + t2*t2, every op rounded
__device__ __forceinline__ float fdist(float r, float g, float b, const float* c) {
    float t0 = __fsub_rn(r, c[0]), t1 = __fsub_rn(g, c[1]), t2 = __fsub_rn(b, c[2]);
    float d = __fmul_rn(t0, t0);
    d = __fadd_rn(d, __fmul_rn(t1, t1));
    d = __fadd_rn(d, __fmul_rn(t2, t2));
    return d;
}

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* sh) {
    v = warp_sum_u64(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < KW; ++i) t += sh[i];
    return t;
}

__device__ __forceinline__ double block_sum_f64(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
#pragma unroll
    for (int i = 0; i < KW; ++i) t += sh[i];
    return t;
}

__global__ void __launch_bounds__(KT) k_kmeans(KmParams P) {
    const int att = blockIdx.x, img = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int U = min(P.count[img], P.max_unique);
    const int K = min(P.k, U);
    const size_t slot = (size_t)img * P.attempts + att;
    if (K <= 1 || P.count[img] > P.max_unique) {   // nothing to cluster / list truncated (see k_kmeans_pick)
        if (tid == 0) {
            P.compact[slot] = 0.0;
            P.iters[slot] = 0;
            P.inexact[slot] = 0;
        }
        return;
    }
    const uint32_t* keys = P.keys + (size_t)img * P.max_unique;
    const uint32_t* wts = P.weights ? P.weights + (size_t)img * P.max_unique : nullptr;
    uint32_t* dist = P.dist + slot * 2 * (size_t)P.max_unique;
    uint8_t* labels = P.labels + slot * P.max_unique;

    __shared__ float s_c[KMAX][3];      // current centres
    __shared__ float s_old[KMAX][3];
    __shared__ unsigned long long s_acc[KW][KMAX][4];  // warp-private sums R,G,B,count
    __shared__ unsigned long long s_tot[KMAX][4];
    __shared__ unsigned long long s_red[KW];
    __shared__ double s_redd[KW];
    __shared__ int s_ci;
    __shared__ double s_p;
    __shared__ int s_flag;
    __shared__ unsigned long long s_far;

    if (tid == 0) s_flag = 0;
    int it;
    if (P.init) {
        for (int i = tid; i < K * 3; i += KT) (&s_c[0][0])[i] = P.init[(size_t)img * P.k * 3 + i];
        for (int i = tid; i < KMAX * 3; i += KT) (&s_old[0][0])[i] = 0.f;
        it = 0;
        __syncthreads();
    } else {
        // ---------------- kmeans++ (cv::generateCentersPP, 3 trials) ----------------
        unsigned long long rs = 0;
        if (tid == 0) {
            rs = P.rng_state[img];
            if (rs == 0) rs = 0xffffffffull;
            const int per_attempt = 1 + 6 * (K - 1);
            for (int i = 0; i < att * per_attempt; ++i) rng_next(rs);
            s_ci = (int)(rng_next(rs) % (uint32_t)U);
        }
        __syncthreads();
        int c0 = s_ci;
        uint32_t ckey = keys[c0];
        if (tid == 0) unpack(ckey, s_c[0][0], s_c[0][1], s_c[0][2]);
        unsigned long long part = 0;
        for (int i = tid; i < U; i += KT) {
            uint32_t d = idist(keys[i], ckey);
            dist[i] = d;
            part += d;
        }
        unsigned long long sum0 = block_sum_u64(part, s_red);
        // contiguous segments per thread for the prefix search
        const int seg = (U + KT - 1) / KT;
        const int lo = min(tid * seg, U), hi = min(lo + seg, U);
        for (int k = 1; k < K; ++k) {
            unsigned long long best_s = ~0ull;
            int best_c = -1;
            for (int trial = 0; trial < 3; ++trial) {
                if (tid == 0) {
                    uint32_t t = rng_next(rs);
                    unsigned long long v = ((unsigned long long)t << 32) | rng_next(rs);
                    s_p = __dmul_rn(__dmul_rn((double)v, 5.4210108624275221700372640043497e-20), (double)sum0);
                    s_ci = U - 1;
                }
                // exclusive prefix of segment sums
                unsigned long long segsum = 0;
                for (int i = lo; i < hi; ++i) segsum += dist[i];
                unsigned long long inc = segsum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    unsigned long long nn = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += nn;
                }
                __syncthreads();  // s_p, s_ci visible; s_red free
                if (lane == 31) s_red[warp] = inc;
                __syncthreads();
                unsigned long long before = inc - segsum;
                for (int i = 0; i < warp; ++i) before += s_red[i];
                const double p = s_p;
                // first i (i <= U-2) with prefix_incl(i) >= p, i.e. the sequential "p -= d; if (p <= 0) break"
                if ((double)(before + segsum) >= p && !((double)before >= p && lo > 0)) {
                    unsigned long long run = before;
                    for (int i = lo; i < hi; ++i) {
                        run += dist[i];
                        if ((double)run >= p) {
                            if (i < U - 1) atomicMin(&s_ci, i);
                            break;
                        }
                    }
                }
                __syncthreads();
                const int ci = s_ci;
                const uint32_t tk = keys[ci];
                unsigned long long ps = 0;
                for (int i = tid; i < U; i += KT) ps += min(idist(keys[i], tk), dist[i]);
                unsigned long long s = block_sum_u64(ps, s_red);
                if (s < best_s) {
                    best_s = s;
                    best_c = ci;
                }
                __syncthreads();
            }
            const uint32_t bk = keys[best_c];
            if (tid == 0) unpack(bk, s_c[k][0], s_c[k][1], s_c[k][2]);
            for (int i = tid; i < U; i += KT) dist[i] = min(idist(keys[i], bk), dist[i]);
            sum0 = best_s;
            __syncthreads();
        }
        it = 1;
    }

    // ------------------------------ Lloyd ------------------------------------------
    for (;;) {
        // assignment to s_c + per-cluster sums
        for (int i = tid; i < KW * KMAX * 4; i += KT) (&s_acc[0][0][0])[i] = 0ull;
        __syncthreads();
        for (int i = tid; i < U; i += KT) {
            uint32_t key = keys[i];
            float r, g, b;
            unpack(key, r, g, b);
            float bd = fdist(r, g, b, s_c[0]);
            int bl = 0;
            for (int k = 1; k < K; ++k) {
                float d = fdist(r, g, b, s_c[k]);
                if (d < bd) {
                    bd = d;
                    bl = k;
                }
            }
            labels[i] = (uint8_t)bl;
            unsigned long long wt = wts ? wts[i] : 1ull;
            atomicAdd(&s_acc[warp][bl][0], (unsigned long long)(key >> 16) * wt);
            atomicAdd(&s_acc[warp][bl][1], (unsigned long long)((key >> 8) & 255u) * wt);
            atomicAdd(&s_acc[warp][bl][2], (unsigned long long)(key & 255u) * wt);
            atomicAdd(&s_acc[warp][bl][3], wt);
        }
        __syncthreads();
        for (int i = tid; i < K * 4; i += KT) {
            unsigned long long t = 0;
            for (int w = 0; w < KW; ++w) t += s_acc[w][i >> 2][i & 3];
            s_tot[i >> 2][i & 3] = t;
        }
        __syncthreads();
        // empty-cluster repair (cv2: biggest cluster gives up its farthest member, last max wins)
        for (int k = 0; k < K; ++k) {
            if (s_tot[k][3] != 0) continue;  // uniform (shared memory)
            int mk = 0;
            for (int k1 = 1; k1 < K; ++k1)
                if (s_tot[mk][3] < s_tot[k1][3]) mk = k1;
            float base[3];
            if (P.exact_sums) {
                for (int j = 0; j < 3; ++j) base[j] = (float)((double)s_tot[mk][j] / (double)s_tot[mk][3]);
            } else {
                float sc = __fdiv_rn(1.f, (float)s_tot[mk][3]);
                for (int j = 0; j < 3; ++j) base[j] = __fmul_rn((float)s_tot[mk][j], sc);
            }
            if (tid == 0) s_far = 0ull;
            __syncthreads();
            unsigned long long best = 0ull;
            bool have = false;
            for (int i = tid; i < U; i += KT) {
                if (labels[i] != mk) continue;
                float r, g, b;
                unpack(keys[i], r, g, b);
                float d = fdist(r, g, b, base);
                unsigned long long cand = ((unsigned long long)__float_as_uint(d) << 32) | (uint32_t)i;
                if (!have || cand > best) best = cand, have = true;
            }
            if (have) atomicMax(&s_far, best + 1ull);  // +1 so that a real candidate beats the 0 sentinel
            __syncthreads();
            if (tid == 0) {
                int far = (int)(uint32_t)((s_far - 1ull) & 0xffffffffull);
                uint32_t fk = keys[far];
                if (!wts) labels[far] = (uint8_t)k;
                // OpenCV stores the donor's provisional mean in old_centers[max_k]; the shift test sees it
                for (int j = 0; j < 3; ++j) s_c[mk][j] = base[j];
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
        // new centres + shift
        if (tid < K) {
            for (int j = 0; j < 3; ++j) {
                s_old[tid][j] = s_c[tid][j];
                float c;
                if (P.exact_sums) {
                    c = (float)((double)s_tot[tid][j] / (double)s_tot[tid][3]);
                } else {
                    if (s_tot[tid][j] >= (1ull << 24)) atomicOr(&s_flag, 1);
                    c = __fmul_rn((float)s_tot[tid][j], __fdiv_rn(1.f, (float)s_tot[tid][3]));
                }
                s_c[tid][j] = c;
            }
        }
        __syncthreads();
        double shift = 0.0;
        const bool first_seeded = (P.init != nullptr && it == 0);
        for (int k = 0; k < K; ++k) {
            double s = 0.0;
            for (int j = 0; j < 3; ++j) {
                double t = (double)__fsub_rn(s_c[k][j], s_old[k][j]);  // float subtraction, as OpenCV
                s = __dadd_rn(s, __dmul_rn(t, t));
            }
            shift = fmax(shift, s);
        }
        ++it;
        const int last_it = P.max_iter > 2 ? P.max_iter : 2;
        const bool last = (it == last_it) || (!first_seeded && shift <= P.eps2);
        if (last) break;
        __syncthreads();
    }
    // compactness with the final centres and the labels of the last assignment
    double part = 0.0;
    for (int i = tid; i < U; i += KT) {
        float r, g, b;
        unpack(keys[i], r, g, b);
        double d = (double)fdist(r, g, b, s_c[labels[i]]);
        part += wts ? d * (double)wts[i] : d;
    }
    double comp = block_sum_f64(part, s_redd);
    if (tid == 0) {
        P.compact[slot] = comp;
        P.iters[slot] = it;
        P.inexact[slot] = s_flag;
    }
    for (int i = tid; i < K * 3; i += KT) P.centers[slot * KMAX * 3 + i] = (&s_c[0][0])[i];
    if (P.sums)
        for (int i = tid; i < K * 4; i += KT) P.sums[slot * KMAX * 4 + i] = s_tot[i >> 2][i & 3];
}

// keep the attempt with the smallest compactness (strict '<': the first wins ties)
__global__ void __launch_bounds__(256) k_kmeans_pick(KmParams P, float* out_centers, int32_t* out_labels,
                                                     double* out_compact, int32_t* out_kused, int32_t* out_iters,
                                                     unsigned long long* out_sums, int32_t* out_sizes,
                                                     int32_t* out_status) {
    const int img = blockIdx.x, tid = threadIdx.x;
    const int U = min(P.count[img], P.max_unique);
    const int K = min(P.k, U);
    const uint32_t* keys = P.keys + (size_t)img * P.max_unique;
    if (P.count[img] > P.max_unique) {
        // the caller's list is too short for this image: clustering a truncated list would silently give a wrong
        // palette, so nothing is clustered; k_used = -1 tells the caller to come back with a list of `count` entries
        if (tid == 0) {
            if (out_kused) out_kused[img] = -1;
            if (out_compact) out_compact[img] = 0.0;
            if (out_iters) out_iters[img] = 0;
            if (out_status) out_status[img] = LLFE_KMEANS_TRUNCATED;
            if (out_sizes)
                for (int i = 0; i < P.k; ++i) out_sizes[(size_t)img * P.k + i] = 0;
        }
        return;
    }
    if (K <= 1) {
        // color_extractor.py:185-186: centres = the unique colours themselves, labels = 0
        if (tid == 0) {
            if (out_kused) out_kused[img] = U > 0 ? 1 : 0;
            if (out_compact) out_compact[img] = 0.0;
            if (out_iters) out_iters[img] = 0;
            if (out_status) out_status[img] = 0;
            if (U > 0) unpack(keys[0], out_centers[(size_t)img * P.k * 3], out_centers[(size_t)img * P.k * 3 + 1],
                              out_centers[(size_t)img * P.k * 3 + 2]);
            if (out_sizes)
                for (int i = 0; i < P.k; ++i) out_sizes[(size_t)img * P.k + i] = (i == 0) ? U : 0;
        }
        if (out_labels)
            for (int i = tid; i < U; i += 256) out_labels[(size_t)img * P.max_unique + i] = 0;
        return;
    }
    int best = 0;
    double bc = P.compact[(size_t)img * P.attempts];
    for (int a = 1; a < P.attempts; ++a) {
        double c = P.compact[(size_t)img * P.attempts + a];
        if (c < bc) {
            bc = c;
            best = a;
        }
    }
    const size_t slot = (size_t)img * P.attempts + best;
    for (int i = tid; i < K * 3; i += 256) out_centers[(size_t)img * P.k * 3 + i] = P.centers[slot * KMAX * 3 + i];
    if (out_labels)
        for (int i = tid; i < U; i += 256) out_labels[(size_t)img * P.max_unique + i] = P.labels[slot * P.max_unique + i];
    if (out_sums && P.sums)
        for (int i = tid; i < K * 4; i += 256) out_sums[(size_t)img * P.k * 4 + i] = P.sums[slot * KMAX * 4 + i];
    if (out_sizes)
        for (int i = tid; i < P.k; i += 256)
            out_sizes[(size_t)img * P.k + i] = i < K ? (int32_t)P.sums[slot * KMAX * 4 + 4 * i + 3] : 0;
    if (tid == 0) {
        if (out_kused) out_kused[img] = K;
        if (out_compact) out_compact[img] = bc;
        if (out_iters) out_iters[img] = P.iters[slot];
        if (out_status) {
            int st = 0;
            for (int a = 0; a < P.attempts; ++a) st |= P.inexact[(size_t)img * P.attempts + a];
            out_status[img] = st;
        }
    }
}

int run_kmeans(llfe_ctx* ctx, KmParams P, int n, float* d_centers, int32_t* d_labels, double* d_compact, int32_t* d_kused,
               int32_t* d_iters, uint64_t* d_sums, int32_t* d_sizes, int32_t* d_status) {
    const size_t slots = (size_t)n * P.attempts;
    const bool fast = P.weights == nullptr && !P.exact_sums;
    // the fast path keeps the lists in shared memory; its global scratch (lists that do not fit) is
    // sized for a bounded number of images per launch
    int dist_images = n;
    if (fast) {
        const size_t per_img = (size_t)P.attempts * P.max_unique * 8;
        dist_images = (int)((size_t)(256u << 20) / per_img);
        if (dist_images < 1) dist_images = 1;
        if (dist_images > n) dist_images = n;
    }
    const size_t dslots = (size_t)dist_images * P.attempts;
    const size_t need = WsCarver::need(dslots * P.max_unique * 8) + WsCarver::need(slots * P.max_unique) +
                        WsCarver::need(slots * KMAX * 3 * 4) + 3 * WsCarver::need(slots * 8) +
                        WsCarver::need(slots * KMAX * 4 * 8) + WsCarver::need((size_t)n * 4);
    void* ws;
    LLFE_TRY(llfe_workspace(ctx, need, &ws));
    WsCarver c(ws);
    P.dist = c.take<uint32_t>(dslots * 2 * P.max_unique);
    P.dist_images = dist_images;
    P.labels = c.take<uint8_t>(slots * P.max_unique);
    P.centers = c.take<float>(slots * KMAX * 3);
    P.compact = c.take<double>(slots);
    P.iters = (int32_t*)c.take<double>(slots);
    P.inexact = (int32_t*)c.take<double>(slots);
    P.sums = (unsigned long long*)c.take<unsigned long long>(slots * KMAX * 4);
    P.order = nullptr;
    if (fast) {
        P.order = c.take<int32_t>((size_t)n);   // filled by launch_kmeans_fast
        LLFE_TRY(launch_kmeans_fast(ctx, P, n));
    } else {
        LLFE_KERNEL(ctx, "k_kmeans");
        k_kmeans<<<dim3(P.attempts, n), KT, 0, ctx->stream>>>(P);
        LLFE_LAUNCHED(ctx);
    }
    LLFE_KERNEL(ctx, "k_kmeans_pick");
    k_kmeans_pick<<<n, 256, 0, ctx->stream>>>(P, d_centers, d_labels, d_compact, d_kused, d_iters,
                                              (unsigned long long*)d_sums, d_sizes, d_status);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

}  // namespace

extern "C" int llfe_kmeans_unique(llfe_ctx* ctx, const uint32_t* d_keys, const int32_t* d_count, int n, int max_unique,
                                  int k, int attempts, int max_iter, double eps, const uint64_t* d_rng_state,
                                  float* d_centers, int32_t* d_labels, double* d_compactness, int32_t* d_k_used,
                                  int32_t* d_cluster_sizes, int32_t* d_status) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_keys != nullptr && d_count != nullptr && d_rng_state != nullptr &&
                   d_centers != nullptr);
    LLFE_CHECK_ARG(n >= 0 && n <= 65535 && max_unique > 0 && k >= 1 && k <= KMAX && attempts >= 1 && attempts <= 64 &&
                   max_iter >= 1);
    if (n == 0) return LLFE_OK;
    KmParams P{};
    P.keys = d_keys;
    P.weights = nullptr;
    P.count = d_count;
    P.max_unique = max_unique;
    P.k = k;
    P.attempts = attempts;
    P.max_iter = max_iter < 2 ? 2 : (max_iter > 100 ? 100 : max_iter);   // cv::kmeans clamps maxCount to [2, 100]
    P.eps2 = eps * eps;
    P.exact_sums = 0;
    P.rng_state = d_rng_state;
    P.init = nullptr;
    // bound the scratch (labels of every attempt): process the batch in chunks of images
    const size_t per_img = (size_t)attempts * ((size_t)max_unique + 2048) + 4096;
    int chunk = (int)((size_t)(512u << 20) / per_img);
    if (chunk < 1) chunk = 1;
    for (int i0 = 0; i0 < n; i0 += chunk) {
        const int m = (n - i0) < chunk ? (n - i0) : chunk;
        KmParams Q = P;
        Q.keys = d_keys + (size_t)i0 * max_unique;
        Q.count = d_count + i0;
        Q.rng_state = d_rng_state + i0;
        LLFE_TRY(run_kmeans(ctx, Q, m, d_centers + (size_t)i0 * k * 3, d_labels ? d_labels + (size_t)i0 * max_unique : nullptr,
                            d_compactness ? d_compactness + i0 : nullptr, d_k_used ? d_k_used + i0 : nullptr, nullptr,
                            nullptr, d_cluster_sizes ? d_cluster_sizes + (size_t)i0 * k : nullptr,
                            d_status ? d_status + i0 : nullptr));
    }
    return LLFE_OK;
}

extern "C" int llfe_kmeans_lloyd(llfe_ctx* ctx, const uint32_t* d_keys, const uint32_t* d_weights, const int32_t* d_count,
                                 int n, int max_unique, int k, int max_iter, double eps, int exact_sums,
                                 const float* d_init_centers, float* d_centers, int32_t* d_labels, int32_t* d_iters,
                                 uint64_t* d_sums_counts) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_keys != nullptr && d_count != nullptr && d_init_centers != nullptr &&
                   d_centers != nullptr);
    LLFE_CHECK_ARG(n >= 0 && n <= 65535 && max_unique > 0 && k >= 1 && k <= KMAX && max_iter >= 1);
    if (n == 0) return LLFE_OK;
    KmParams P{};
    P.keys = d_keys;
    P.weights = d_weights;
    P.count = d_count;
    P.max_unique = max_unique;
    P.k = k;
    P.attempts = 1;
    // cv2's rule (exact_sums = 0) inherits cv::kmeans' clamp of maxCount to [2, 100]; the pinned exact-sum rule
    // of the per-pixel mode keeps the caller's limit
    P.max_iter = exact_sums ? max_iter : (max_iter < 2 ? 2 : (max_iter > 100 ? 100 : max_iter));
    P.eps2 = eps * eps;
    P.exact_sums = exact_sums ? 1 : 0;
    P.rng_state = nullptr;
    P.init = d_init_centers;
    const size_t per_img = (size_t)max_unique * 9 + 4096;
    int chunk = (int)((size_t)(256u << 20) / per_img);
    if (chunk < 1) chunk = 1;
    for (int i0 = 0; i0 < n; i0 += chunk) {
        const int m = (n - i0) < chunk ? (n - i0) : chunk;
        KmParams Q = P;
        Q.keys = d_keys + (size_t)i0 * max_unique;
        Q.weights = d_weights ? d_weights + (size_t)i0 * max_unique : nullptr;
        Q.count = d_count + i0;
        Q.init = d_init_centers + (size_t)i0 * k * 3;
        LLFE_TRY(run_kmeans(ctx, Q, m, d_centers + (size_t)i0 * k * 3, d_labels ? d_labels + (size_t)i0 * max_unique : nullptr,
                            nullptr, nullptr, d_iters ? d_iters + i0 : nullptr,
                            d_sums_counts ? d_sums_counts + (size_t)i0 * k * 4 : nullptr, nullptr, nullptr));
    }
    return LLFE_OK;
}

