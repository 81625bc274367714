// Declarations shared by the two k-means kernels (k_kmeans.cu: general path with weights /
// exact sums; k_kmeans_fast.cu: shared-memory-resident path for unweighted lists).
#pragma once
#include "llfe_common.cuh"
#include "llfe_device.cuh"

constexpr int KMAX = 32;

struct KmParams {
    const uint32_t* keys;      // [n][max_unique]  R<<16|G<<8|B
    const uint32_t* weights;   // [n][max_unique] or null
    const int32_t* count;      // [n]
    int max_unique, k, attempts, max_iter;
    double eps2;
    int exact_sums;            // 1: c = float(double(sum)/double(cnt))
    const uint64_t* rng_state; // [n] (PP mode) or null
    const float* init;         // [n][k][3] (seeded mode) or null
    // scratch, per (image, attempt)
    uint32_t* dist;            // [dist_images][attempts][2 * max_unique]
    int dist_images;           // images the dist scratch holds (fast path: one launch of the global variant)
    uint8_t* labels;           // [n][attempts][max_unique]
    float* centers;            // [n][attempts][KMAX][3]
    double* compact;           // [n][attempts]
    int32_t* iters;            // [n][attempts]
    int32_t* inexact;          // [n][attempts]
    unsigned long long* sums;  // [n][attempts][KMAX][4]  final sums/counts
    unsigned long long* dbg;   // optional [n][attempts][8] phase clocks (LLFE_KMEANS_DEBUG), else null
    const int32_t* order;      // optional [n]: image handled by blockIdx.y (longest lists first), else null = identity
};

__device__ __forceinline__ uint32_t idist(uint32_t a, uint32_t b) {
    int dr = (int)((a >> 16) & 255u) - (int)((b >> 16) & 255u);
    int dg = (int)((a >> 8) & 255u) - (int)((b >> 8) & 255u);
    int db = (int)(a & 255u) - (int)(b & 255u);
    return (uint32_t)(dr * dr + dg * dg + db * db);
}

__device__ __forceinline__ uint32_t rng_next(unsigned long long& s) {
    s = (unsigned long long)(uint32_t)s * 4164903690ull + (s >> 32);
    return (uint32_t)s;
}

int launch_kmeans_fast(llfe_ctx* ctx, const KmParams& P, int n);
