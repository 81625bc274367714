// Canny hysteresis + 3x3 dilate + bit-plane -> u8 mask in ONE launch, one thread-block
// CLUSTER per image (cv2.Canny's hysteresis stage and cv2.dilate of shape_analyzer pyc L24-28).
//
//   * the image's rows are split into CL strips, one CTA each.  A CTA keeps its strip of the
//     weak plane and of the edge plane in shared memory, EXTENDED by HC_D + 1 rows of its
//     neighbours on both sides: the HC_D nearest rows are worked on by both CTAs (so a chain that
//     wiggles across a strip boundary is resolved inside one CTA), the outermost row is context;
//   * inside a CTA every WARP owns a range of rows (ranges overlap by HC_D rows for the same
//     reason).  One step handles a whole row: lane L holds WPL consecutive plane words (32 lanes
//     cover the image width), pulls in the rows above/below (3x3 neighbourhood on bit level) and
//     closes the row horizontally with fill_word + warp-shuffle carries until nothing moves.
//     The warp walks DOWN its live rows (rows that still have weak-but-not-edge pixels) and steps
//     BACK one row whenever a row gained pixels, so chains running down, up or zig-zag inside
//     the range are resolved by one walk, not by one Jacobi iteration per pixel;
//   * ranges iterate (block barrier) until the CTA is stable, then strips OR-merge the rows they
//     share through DISTRIBUTED SHARED MEMORY and repeat while any merge changed something
//     (cluster barrier, a flag in rank 0's shared memory);
//   * when nothing changes any more every CTA expands its own rows to the u8 mask (optionally
//     3x3-dilated), so the edge plane never goes back to HBM.
//
// HBM traffic: read 2 * P/8 (planes), write P (mask).  The fallback for images whose strips do
// not fit in shared memory is the multi-launch strip kernel of k_canny.cu.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "llfe_common.cuh"
#include "llfe_device.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int HC_THREADS = 256;
constexpr int HC_NW = HC_THREADS / 32;
constexpr int HC_D = 2;                 // rows shared with each neighbour (strip and warp range)
constexpr unsigned FULLM = 0xffffffffu;

struct HcArgs {
    const uint32_t* weak;    // [n][h][wpr]
    const uint32_t* strong;  // [n][h][wpr]
    uint8_t* mask;           // [n][h][w]
    int h, w, wpr, rps;      // rps = rows per strip
    int aligned;             // 16-byte stores allowed
    unsigned long long* dbg; // optional [n][CL][8] phase clocks / counters (LLFE_HYST_DEBUG), else null
};

template <int WPL>
struct Words {
    uint32_t v[WPL];
};

template <int WPL>
__device__ __forceinline__ Words<WPL> lds_words(const uint32_t* p) {
    Words<WPL> r;
    if (WPL == 1) {
        r.v[0] = p[0];
    } else if (WPL == 2) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        r.v[0] = t.x;
        r.v[1] = t.y;
    } else {
#pragma unroll
        for (int j = 0; j < WPL; j += 4) {
            const uint4 t = *reinterpret_cast<const uint4*>(p + j);
            r.v[j] = t.x;
            r.v[j + 1] = t.y;
            r.v[j + 2] = t.z;
            r.v[j + 3] = t.w;
        }
    }
    return r;
}

// 3x3 neighbourhood on bit level: v = OR of the rows above, at and below; returns w & spread(v) & ~e per word
template <int WPL>
__device__ __forceinline__ uint32_t hc_grow(const uint32_t* v, const Words<WPL>& w, const Words<WPL>& e, Words<WPL>& ne,
                                            int lane) {
    uint32_t vl = __shfl_up_sync(FULLM, v[WPL - 1], 1), vr = __shfl_down_sync(FULLM, v[0], 1);
    if (lane == 0) vl = 0u;
    if (lane == 31) vr = 0u;
    uint32_t grow = 0u;
#pragma unroll
    for (int j = 0; j < WPL; ++j) {
        const uint32_t pl = j > 0 ? v[j - 1] : vl, pr = j + 1 < WPL ? v[j + 1] : vr;
        const uint32_t spread = v[j] | (v[j] << 1) | (v[j] >> 1) | (pl >> 31) | (pr << 31);
        const uint32_t g = w.v[j] & spread & ~e.v[j];
        grow |= g;
        ne.v[j] = e.v[j] | g;
    }
    return grow;
}

// horizontal closure of a row: fill inside the words, carry across word / lane boundaries until nothing moves
template <int WPL>
__device__ __forceinline__ void hc_close(Words<WPL>& ne, const Words<WPL>& w, int lane) {
    for (;;) {
#pragma unroll
        for (int j = 0; j < WPL; ++j) ne.v[j] = fill_word(ne.v[j], w.v[j]);
        uint32_t cl = __shfl_up_sync(FULLM, ne.v[WPL - 1], 1), cr = __shfl_down_sync(FULLM, ne.v[0], 1);
        if (lane == 0) cl = 0u;
        if (lane == 31) cr = 0u;
        uint32_t add_any = 0u;
#pragma unroll
        for (int j = 0; j < WPL; ++j) {
            const uint32_t pl = j > 0 ? ne.v[j - 1] : cl, pr = j + 1 < WPL ? ne.v[j + 1] : cr;
            const uint32_t add = w.v[j] & ~ne.v[j] & ((pl >> 31) | (pr << 31));
            add_any |= add;
            ne.v[j] |= add;
        }
        if (!__any_sync(FULLM, add_any != 0u)) break;
    }
}

// One propagation step of local row i (1 <= i <= L-2), executed by a whole warp.  Returns
// (warp-uniform) 0: the row has no weak-but-not-edge pixel left, 1: nothing gained, 2: gained pixels;
// + 4 when the row ABOVE gained pixels too.  A chain that wiggles between two adjacent rows (the usual
// shape of a weak, nearly horizontal edge) is followed in registers: after row i grew, row i-1 is grown
// from it, then row i from row i-1, ... until neither moves, and both rows are written back once.
template <int WPL>
__device__ __forceinline__ int hc_row_step(uint32_t* E, const uint32_t* W, int i, int lane) {
    constexpr int SP = 32 * WPL;
    uint32_t* erow = E + i * SP + lane * WPL;
    const Words<WPL> e = lds_words<WPL>(erow);
    const Words<WPL> w = lds_words<WPL>(W + i * SP + lane * WPL);
    const Words<WPL> up = lds_words<WPL>(erow - SP);
    const Words<WPL> dn = lds_words<WPL>(erow + SP);
    uint32_t cand = 0u;
    uint32_t v[WPL];
#pragma unroll
    for (int j = 0; j < WPL; ++j) {
        cand |= w.v[j] & ~e.v[j];
        v[j] = up.v[j] | e.v[j] | dn.v[j];
    }
    Words<WPL> ne;
    const uint32_t grow = hc_grow<WPL>(v, w, e, ne, lane);
    const uint32_t bc = __ballot_sync(FULLM, cand != 0u), bg = __ballot_sync(FULLM, grow != 0u);
    if (bg == 0u) return bc ? 1 : 0;
    hc_close<WPL>(ne, w, lane);
    int above_changed = 0;
    if (i >= 2) {
        // ping-pong with the row above, in registers
        const Words<WPL> w1 = lds_words<WPL>(W + (i - 1) * SP + lane * WPL);
        const Words<WPL> upup = lds_words<WPL>(erow - 2 * SP);
        Words<WPL> e1 = up;   // row i-1 as loaded
        for (;;) {
            uint32_t v1[WPL];
#pragma unroll
            for (int j = 0; j < WPL; ++j) v1[j] = upup.v[j] | e1.v[j] | ne.v[j];
            Words<WPL> n1;
            const uint32_t g1 = hc_grow<WPL>(v1, w1, e1, n1, lane);
            if (!__any_sync(FULLM, g1 != 0u)) break;
            hc_close<WPL>(n1, w1, lane);
            e1 = n1;
            above_changed = 4;
            uint32_t v0[WPL];
#pragma unroll
            for (int j = 0; j < WPL; ++j) v0[j] = e1.v[j] | ne.v[j] | dn.v[j];
            Words<WPL> n0;
            const uint32_t g0 = hc_grow<WPL>(v0, w, ne, n0, lane);
            if (!__any_sync(FULLM, g0 != 0u)) break;
            hc_close<WPL>(n0, w, lane);
            ne = n0;
        }
        if (above_changed) {
#pragma unroll
            for (int j = 0; j < WPL; ++j)
                if (e1.v[j] != up.v[j]) atomicOr(erow - SP + j, e1.v[j]);
        }
    }
    // rows in the overlap of two ranges can be written by two warps: OR, never overwrite
#pragma unroll
    for (int j = 0; j < WPL; ++j)
        if (ne.v[j] != e.v[j]) atomicOr(erow + j, ne.v[j]);
    return 2 | above_changed;
}

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t smem_addr, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(g) : "memory");
}

template <int CL, int WPL, bool DILATE>
__global__ void __launch_bounds__(HC_THREADS) k_hyst_mask(HcArgs A) {
    extern __shared__ __align__(16) uint32_t sm[];
    __shared__ uint32_t s_flag[2];
    __shared__ int s_live;
    constexpr int SP = 32 * WPL;
    constexpr int D = HC_D;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();  // == blockIdx.x
    const int img = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wpr = A.wpr, rps = A.rps, h = A.h;
    const int r0 = rank * rps;
    const int rows = max(0, min(rps, h - r0));
    const int L = rps + 2 * (D + 1);       // local rows: global rows g0 .. g0 + L - 1
    const int g0 = r0 - (D + 1);
    uint32_t* E = sm;                      // [L][SP]
    uint32_t* W = sm + L * SP;             // [L][SP]  (context rows 0 and L-1 stay zero: never grown)
    uint32_t* chg = sm + 2 * L * SP;       // [L + 2] epoch of the last change of each row (index i + 1; 0 = never)
    const size_t plane = (size_t)h * wpr;
    const uint32_t* gw = A.weak + img * plane;
    const uint32_t* gs = A.strong + img * plane;

    if (tid == 0) {
        s_flag[0] = s_flag[1] = 0u;
        s_live = 0;
    }
    const long long t_start = clock64();
    long long t_local = 0, t_sync = 0;
    unsigned n_iter = 0, n_round = 0;
    // ---- load: zero everything, then stream the in-image rows with cp.async (all loads in flight) ----
    for (int q = tid; q < 2 * L * SP / 4; q += HC_THREADS) reinterpret_cast<uint4*>(sm)[q] = make_uint4(0u, 0u, 0u, 0u);
    for (int q = tid; q < L + 2; q += HC_THREADS) chg[q] = 0u;
    __syncthreads();
    {
        const uint32_t sE = (uint32_t)__cvta_generic_to_shared(E), sW = (uint32_t)__cvta_generic_to_shared(W);
        if ((wpr & 3) == 0) {
            const int cpr = wpr >> 2;  // 16-byte chunks per row
            for (int q = tid; q < L * cpr; q += HC_THREADS) {
                const int i = q / cpr, c4 = (q - i * cpr) * 4, gr = g0 + i;
                if (gr < 0 || gr >= h) continue;
                cp_async16(sE + (uint32_t)(i * SP + c4) * 4u, gs + (size_t)gr * wpr + c4);
                if (i > 0 && i < L - 1) cp_async16(sW + (uint32_t)(i * SP + c4) * 4u, gw + (size_t)gr * wpr + c4);
            }
        } else {
            for (int q = tid; q < L * wpr; q += HC_THREADS) {
                const int i = q / wpr, c = q - i * wpr, gr = g0 + i;
                if (gr < 0 || gr >= h) continue;
                cp_async4(sE + (uint32_t)(i * SP + c) * 4u, gs + (size_t)gr * wpr + c);
                if (i > 0 && i < L - 1) cp_async4(sW + (uint32_t)(i * SP + c) * 4u, gw + (size_t)gr * wpr + c);
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();
    // ---- row range of this warp (processed rows are 1 .. L-2), extended by D rows on both sides ----
    const int lp = L - 2;
    const int rpw = (lp + HC_NW - 1) / HC_NW;
    const int ca = 1 + min(lp, warp * rpw), cb = 1 + min(lp, (warp + 1) * rpw);
    const int ra = max(1, ca - D), rb = min(L - 1, cb + D);   // [ra, rb), at most 64 rows (checked by the launcher)
    unsigned long long live = 0ull;                           // bit (i - ra): row i still has candidates
    for (int i = ra; i < rb; ++i) {
        const Words<WPL> e = lds_words<WPL>(E + i * SP + lane * WPL);
        const Words<WPL> w = lds_words<WPL>(W + i * SP + lane * WPL);
        uint32_t cand = 0u;
#pragma unroll
        for (int j = 0; j < WPL; ++j) cand |= w.v[j] & ~e.v[j];
        if (__any_sync(FULLM, cand != 0u)) live |= 1ull << (i - ra);
    }
    unsigned long long dirty = live;
    uint32_t ep = 0u;
    if (live && lane == 0) s_live = 1;  // benign race: everyone writes 1
    __syncthreads();
    const bool has_live = s_live != 0;
    const long long t_loaded = clock64();

    for (int round = 0;; ++round) {
        const long long ta = clock64();
        // ---- local convergence: walk down the rows that can still grow, step back when a row gained pixels.
        // A row needs another look only if it, or a row next to it, changed since it was last processed:
        // `dirty` tracks that per warp (own steps directly, other warps' steps and merges through the
        // per-row change epochs), so converged parts of the strip are not walked again.
        if (has_live) {
            for (;;) {
                ++n_iter;
                ++ep;
                if (ep > 1) {   // rows next to something that changed in the previous iteration / merge
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int i = ra + 32 * half + lane;
                        bool f = false;
                        if (i < rb) f = chg[i] == ep - 1 || chg[i + 1] == ep - 1 || chg[i + 2] == ep - 1;  // rows i-1, i, i+1
                        const unsigned long long b = __ballot_sync(FULLM, f);
                        dirty |= b << (32 * half);
                    }
                }
                bool changed = false;
                int i = ra;
                while (i < rb) {
                    const unsigned long long m = (live & dirty) >> (i - ra);
                    if (m == 0ull) break;
                    i += __ffsll((long long)m) - 1;            // next row that is live and dirty
                    const int res4 = hc_row_step<WPL>(E, W, i, lane);
                    const int res = res4 & 3;
                    dirty &= ~(1ull << (i - ra));
                    if (res == 0) live &= ~(1ull << (i - ra));
                    if (res == 2) {
                        changed = true;
                        if (lane == 0) {
                            chg[i + 1] = ep;
                            if (res4 & 4) chg[i] = ep;       // the row above grew with it
                        }
                        if (i + 1 < rb) dirty |= 1ull << (i + 1 - ra);
                        // the row above is already consistent with this one; look further up only if it grew
                        if ((res4 & 4) && i - 2 >= ra) {
                            dirty |= 1ull << (i - 2 - ra);
                            if ((live >> (i - 2 - ra)) & 1ull) {
                                i -= 2;
                                continue;
                            }
                        }
                    }
                    ++i;
                }
                if (!__syncthreads_or(changed)) break;
            }
        }
        const long long tb = clock64();
        cluster.sync();  // every strip of the image is locally stable
        // ---- OR-merge the 2 (D + 1) rows shared with each neighbour through distributed shared memory --
        bool mchanged = false;
        constexpr int NS = 2 * (D + 1);
        ++ep;   // the merge is an epoch of its own: rows it changes are looked at by the next walk
        if (rank > 0) {  // my rows 0 .. NS-1  <->  rows rps .. L-1 of the strip above
            const uint32_t* nb = cluster.map_shared_rank(E, rank - 1) + rps * SP;
            for (int q = tid; q < NS * SP; q += HC_THREADS) {
                const uint32_t v = nb[q], mine = E[q];
                if (v & ~mine) {
                    E[q] = mine | v;
                    chg[q / SP + 1] = ep;
                    mchanged = true;
                }
            }
        }
        __syncthreads();      // short strips: the two windows can overlap, keep their read-modify-writes apart
        if (rank + 1 < CL) {  // my rows rps .. L-1  <->  rows 0 .. NS-1 of the strip below
            const uint32_t* nb = cluster.map_shared_rank(E, rank + 1);
            uint32_t* me = E + rps * SP;
            for (int q = tid; q < NS * SP; q += HC_THREADS) {
                const uint32_t v = nb[q], mine = me[q];
                if (v & ~mine) {
                    me[q] = mine | v;
                    chg[rps + q / SP + 1] = ep;
                    mchanged = true;
                }
            }
        }
        mchanged = __syncthreads_or(mchanged) && has_live;
        if (mchanged && tid == 0) atomicOr(cluster.map_shared_rank(&s_flag[round & 1], 0), 1u);
        cluster.sync();  // flags of this round are complete; nobody still reads my rows
        const uint32_t again = *(volatile uint32_t*)cluster.map_shared_rank(&s_flag[round & 1], 0);
        if (rank == 0 && tid == 0) s_flag[(round + 1) & 1] = 0u;
        t_local += tb - ta;
        t_sync += clock64() - tb;
        ++n_round;
        if (!again) break;
    }
    // A round ends only when no merge changed a strip that can still grow, so every copy of a shared
    // row is final.  No CTA may exit while a neighbour can still read its shared memory (`again`).
    cluster.sync();
    const long long t_conv = clock64();

    // ---- expand (and dilate) the strip's own rows to the u8 mask ----------------------------------------
    uint8_t* m = A.mask + (size_t)img * h * A.w;
    const int halves = 2 * wpr;
    for (int r = warp; r < rows; r += HC_NW) {
        const uint32_t* row = E + (r + D + 1) * SP;
        const int y = r0 + r;
        for (int idx = lane; idx < halves; idx += 32) {
            const int c = idx >> 1, half = idx & 1;
            uint32_t v;
            if (DILATE) {
                const uint32_t ctr = row[c - SP] | row[c] | row[c + SP];
                uint32_t lft = 0u, rgt = 0u;
                if (c > 0) lft = row[c - 1 - SP] | row[c - 1] | row[c - 1 + SP];
                if (c + 1 < wpr) rgt = row[c + 1 - SP] | row[c + 1] | row[c + 1 + SP];
                v = ctr | (ctr << 1) | (ctr >> 1) | (lft >> 31) | (rgt << 31);
            } else {
                v = row[c];
            }
            const uint32_t bits = (v >> (16 * half)) & 0xffffu;
            const int x = c * 32 + half * 16;
            if (x >= A.w) continue;
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t nib = (bits >> (4 * k)) & 0xfu;
                o[k] = ((nib * 0x00204081u) & 0x01010101u) * 255u;  // bit i -> byte i
            }
            uint8_t* dst = m + (size_t)y * A.w + x;
            if (A.aligned && x + 16 <= A.w) {
                __stcs(reinterpret_cast<uint4*>(dst), make_uint4(o[0], o[1], o[2], o[3]));
            } else {
                for (int k = 0; k < 16 && x + k < A.w; ++k) dst[k] = (uint8_t)(o[k >> 2] >> (8 * (k & 3)));
            }
        }
    }
    if (A.dbg && tid == 0) {
        unsigned long long* d = A.dbg + ((size_t)img * CL + rank) * 8;
        d[0] = (unsigned long long)(t_loaded - t_start);
        d[1] = (unsigned long long)t_local;
        d[2] = (unsigned long long)t_sync;
        d[3] = (unsigned long long)(t_conv - t_loaded);
        d[4] = (unsigned long long)(clock64() - t_conv);
        d[5] = n_iter;
        d[6] = n_round;
        d[7] = has_live;
    }
}

// shared memory of one CTA; 0 when the strip does not fit the kernel's limits
size_t hc_smem(int rps, int sp) {
    const int L = rps + 2 * (HC_D + 1);
    const int rpw = (L - 2 + HC_NW - 1) / HC_NW;
    if (rpw + 2 * HC_D > 64) return (size_t)1 << 40;   // the per-warp live mask is 64 bits
    return ((size_t)2 * L * sp + L + 2 + 2) * sizeof(uint32_t);
}

template <int CL, int WPL, bool DILATE>
int launch_hc(llfe_ctx* ctx, const HcArgs& A, int n, size_t smem) {
    if (llfe_first_use(ctx, (const void*)k_hyst_mask<CL, WPL, DILATE>)) {
        LLFE_CUDA(cudaFuncSetAttribute(k_hyst_mask<CL, WPL, DILATE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(ctx->smem_optin - 1024)));
        if (CL > 8)
            LLFE_CUDA(cudaFuncSetAttribute(k_hyst_mask<CL, WPL, DILATE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL, n, 1);
    cfg.blockDim = dim3(HC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    LLFE_KERNEL(ctx, "k_hyst_mask");
    LLFE_CUDA(cudaLaunchKernelEx(&cfg, k_hyst_mask<CL, WPL, DILATE>, A));
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

template <int WPL>
int launch_hc_w(llfe_ctx* ctx, HcArgs A, int n, int dilate) {
    const size_t budget = ctx->smem_optin - 1024;
    const int sp = 32 * WPL;
    const int rps8 = ceil_div(A.h, 8) < 1 ? 1 : ceil_div(A.h, 8);
    if (hc_smem(rps8, sp) <= budget) {
        A.rps = rps8;
        return dilate ? launch_hc<8, WPL, true>(ctx, A, n, hc_smem(rps8, sp)) : launch_hc<8, WPL, false>(ctx, A, n, hc_smem(rps8, sp));
    }
    const int rps16 = ceil_div(A.h, 16);
    if (hc_smem(rps16, sp) <= budget) {
        A.rps = rps16;
        return dilate ? launch_hc<16, WPL, true>(ctx, A, n, hc_smem(rps16, sp))
                      : launch_hc<16, WPL, false>(ctx, A, n, hc_smem(rps16, sp));
    }
    return LLFE_E_UNSUPPORTED;
}

}  // namespace

// weak / strong planes -> final (optionally dilated) u8 edge mask.  Returns LLFE_E_UNSUPPORTED
// without touching anything when the strips do not fit (the caller then takes the strip kernels).
int launch_hysteresis_mask_cluster(llfe_ctx* ctx, const uint32_t* weak, const uint32_t* strong, int n, int h, int w,
                                   int dilate, uint8_t* mask) {
    const int wpr = plane_wpr(w);
    HcArgs A;
    A.weak = weak;
    A.strong = strong;
    A.mask = mask;
    A.h = h;
    A.w = w;
    A.wpr = wpr;
    A.rps = 0;
    A.aligned = (w % 16 == 0) && ((uintptr_t)mask % 16 == 0);
    // phase clocks [n][16][8] u64, only into a buffer registered (and validated) by llfe_set_debug_buffer
    A.dbg = (ctx->dbg_hyst && ctx->dbg_hyst_bytes >= (size_t)n * 16 * 8 * 8) ? ctx->dbg_hyst : nullptr;
    if (wpr <= 32) return launch_hc_w<1>(ctx, A, n, dilate);
    if (wpr <= 64) return launch_hc_w<2>(ctx, A, n, dilate);
    if (wpr <= 128) return launch_hc_w<4>(ctx, A, n, dilate);
    return LLFE_E_UNSUPPORTED;
}
