// Shared by k_pixels.cu and k_colorhist.cu: the fused all-reduce + centre update over peer memory (config 5, N > 1).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "k_kmeans_shared.cuh"

// What crosses NVLink per Lloyd iteration is K x 4 exact u64 sums (512 B at K = 16).  Instead of an ncclAllReduce
// (20-50 us of launch + protocol latency for 512 bytes) followed by the update kernel, ONE CTA per rank stores its
// partial sums straight into every peer's mailbox (P2P stores through NVLink / NVSwitch), raises a flag there, waits for
// the flags of all peers in its own mailbox, adds the G partial sums in rank order (integers: the total does not depend
// on the order) and runs the centre update.  Mailboxes are double-buffered by the parity of an epoch counter: a rank can
// only write epoch e + 2 into a peer after that peer has published e + 1, i.e. after it has finished reading e.  Every
// rank takes the converged / frozen early exit on the same call (identical state everywhere).
constexpr int P2P_MAXW = 16;
struct P2PMailbox {
    unsigned long long slots[2][P2P_MAXW][KMAX * 4];
    unsigned int flags[2][P2P_MAXW];
    unsigned int epoch;
    unsigned int bflags[2][P2P_MAXW];   // the stand-alone barrier (p2p_barrier) has its own flags and epoch
    unsigned int bepoch;
    unsigned int pad[30];
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Called by ALL threads of one CTA (>= 128 threads, K * 4 <= blockDim.x).  peers == nullptr: a world of one, the
// partial sums are the totals.  s_tot: K * 4 words of shared memory.  Returns with `sums` cleared, `totals`, `centers`,
// `state`, `shift_out` updated as llfe_kmeans_update documents; a no-op when the state is converged / frozen.
__device__ __forceinline__ void p2p_exchange_and_update(int K, unsigned long long* sums, P2PMailbox* const* peers, int rank,
                                                        int world, float* centers, int max_iter, double eps2,
                                                        int32_t* state, double* shift_out, unsigned long long* totals,
                                                        unsigned long long* s_tot) {
    const unsigned FULLM = 0xffffffffu;
    const int t = threadIdx.x;
    const int blocked = ((volatile int32_t*)state)[1] | ((volatile int32_t*)state)[3];
    const int it0 = ((volatile int32_t*)state)[0];
    P2PMailbox* mine = peers ? peers[rank] : nullptr;
    const unsigned int e = mine ? mine->epoch + 1u : 0u;
    __syncthreads();
    if (blocked) return;   // converged / frozen: no-op on every rank alike (block-uniform)
    const int p = (int)(e & 1u);
    if (t < K * 4) {
        const unsigned long long v = __ldcg(&sums[t]);
        sums[t] = 0ull;                       // the accumulator is empty again for the next assignment
        if (mine) {
            for (int q = 0; q < world; ++q) peers[q]->slots[p][rank][t] = v;
        } else {
            s_tot[t] = v;
            totals[t] = v;
        }
    }
    if (mine) {
        __threadfence_system();
        __syncthreads();
        if (t < world) {
            st_release_sys(&peers[t]->flags[p][rank], e);
            while (ld_acquire_sys(&mine->flags[p][t]) != e) {
            }
        }
        __syncthreads();
        if (t < K * 4) {
            unsigned long long tot = 0ull;
            for (int q = 0; q < world; ++q) tot += *(volatile unsigned long long*)&mine->slots[p][q][t];
            s_tot[t] = tot;
            totals[t] = tot;
        }
        if (t == 0) mine->epoch = e;
    }
    __syncthreads();
    if (t >= 32) return;
    const int lane = t;
    unsigned long long s[4] = {0ull, 0ull, 0ull, 1ull};
    if (lane < K) {
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] = s_tot[4 * lane + j];
    }
    const int n_empty = __popc(__ballot_sync(FULLM, s[3] == 0ull));
    if (n_empty) {
        if (lane == 0) {
            state[2] = n_empty;
            state[3] = 1;
        }
        return;
    }
    double sh = 0.0;
    if (lane < K) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float c = (float)((double)s[j] / (double)s[3]);
            const double d = (double)__fsub_rn(c, __ldcg(&centers[3 * lane + j]));
            sh = __dadd_rn(sh, __dmul_rn(d, d));
            centers[3 * lane + j] = c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sh = fmax(sh, __shfl_xor_sync(FULLM, sh, o));
    if (lane == 0) {
        const int it = it0 + 1;
        const int last_it = max_iter > 2 ? max_iter : 2;
        state[0] = it;
        state[2] = 0;
        state[1] = (it == last_it) || (it0 > 0 && sh <= eps2);
        if (shift_out) *shift_out = sh;
    }
}

// Cross-rank barrier on the device, called by one CTA of >= `world` threads on every rank: "everything this rank enqueued
// before is visible to the peers, and the peers have got as far".  Same flag protocol as above with its own epoch.
__device__ __forceinline__ void p2p_barrier(P2PMailbox* const* peers, int rank, int world) {
    const int t = threadIdx.x;
    P2PMailbox* mine = peers[rank];
    const unsigned int e = mine->bepoch + 1u;
    __threadfence_system();
    __syncthreads();
    if (t < world) {
        st_release_sys(&peers[t]->bflags[e & 1u][rank], e);
        while (ld_acquire_sys(&mine->bflags[e & 1u][t]) != e) {
        }
    }
    __syncthreads();
    if (t == 0) mine->bepoch = e;
}
