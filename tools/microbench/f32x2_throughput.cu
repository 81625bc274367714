// microbenchmark: scalar FADD/FMUL vs packed add/mul .f32x2 throughput on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float a[8]; unsigned long long p[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f); }
    const float c = seed * 0.999f; const unsigned long long c2 = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i] = __fadd_rn(__fmul_rn(a[i], c), c); }
            else { p[i] = add2(mul2(p[i], c2), c2); }
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += MODE == 0 ? a[i] : __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 1.0001f); else k<1><<<148 * 8, 256>>>(d, iters, 1.0001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)148 * 8 * 256 * iters * 8 * 2 * (mode ? 2 : 1);   // scalar flop-instr results
        printf("mode %d: %.3f ms, %.2f T lane-results/s (instr/thread=%d)\n", mode, ms, ops / ms / 1e9, iters * 16);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
