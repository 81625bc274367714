// Canny edge chain (cv2.Canny(img, low, high), aperture 3, L1 norm) + 3x3 dilate.
//
//   front      : Sobel (BORDER_REPLICATE) -> |dx|+|dy| -> non-max suppression ->
//                two BIT PLANES per image: weak (kept, mag > low) and strong
//                (kept, mag > high).  One bit per pixel: 1/32 of a u8 map.
//   hysteresis : edges = fixed point of  E <- weak & dilate8(E),  E0 = strong,
//                done on the bit planes with word-parallel fills; strips of rows
//                converge locally in shared memory, cross-strip propagation is
//                finished by one CTA per image (no host round trips, no grid sync).
//   expand     : bit plane -> u8 {0,255} mask, optionally 3x3-dilated on the way.
#include "llfe_common.cuh"
#include "llfe_device.cuh"

namespace {

// ------------------------------------------------------------------ front ---
constexpr int FTW = 64;  // tile width  (2 plane words)
constexpr int FTH = 32;  // tile height

__global__ void __launch_bounds__(256) k_canny_front(const uint8_t* __restrict__ gray, int h, int w, int low, int high,
                                                     uint32_t* __restrict__ weak, uint32_t* __restrict__ strong) {
    __shared__ uint8_t g[FTH + 4][FTW + 4 + 4];
    __shared__ uint16_t mg[FTH + 2][FTW + 2 + 2];
    __shared__ short2 dxy[FTH][FTW];
    const int img = blockIdx.z;
    const uint8_t* s = gray + (size_t)img * h * w;
    const int wpr = plane_wpr(w);
    uint32_t* pw = weak + (size_t)img * h * wpr;
    uint32_t* ps = strong + (size_t)img * h * wpr;
    const int x0 = blockIdx.x * FTW, y0 = blockIdx.y * FTH;
    const int tid = threadIdx.x;

    // stage the source with replicated borders: rows y0-2.., cols x0-2..
    for (int i = tid; i < (FTH + 4) * (FTW + 4); i += 256) {
        int ry = i / (FTW + 4), rx = i - ry * (FTW + 4);
        int y = clampi(y0 - 2 + ry, 0, h - 1), x = clampi(x0 - 2 + rx, 0, w - 1);
        g[ry][rx] = s[(size_t)y * w + x];
    }
    __syncthreads();
    // gradient magnitude on the tile + 1 halo; zero outside the image
    for (int i = tid; i < (FTH + 2) * (FTW + 2); i += 256) {
        int ry = i / (FTW + 2), rx = i - ry * (FTW + 2);
        int y = y0 - 1 + ry, x = x0 - 1 + rx;
        int m = 0;
        if (y >= 0 && y < h && x >= 0 && x < w) {
            // g index of (y, x) is [ry + 1][rx + 1]
            const int a = ry + 1, b = rx + 1;
            int tl = g[a - 1][b - 1], tc = g[a - 1][b], tr = g[a - 1][b + 1];
            int ml = g[a][b - 1], mr = g[a][b + 1];
            int bl = g[a + 1][b - 1], bc = g[a + 1][b], br = g[a + 1][b + 1];
            int dx = (tr + 2 * mr + br) - (tl + 2 * ml + bl);
            int dy = (bl + 2 * bc + br) - (tl + 2 * tc + tr);
            m = abs(dx) + abs(dy);
            if (ry >= 1 && ry <= FTH && rx >= 1 && rx <= FTW) dxy[ry - 1][rx - 1] = make_short2((short)dx, (short)dy);
        }
        mg[ry][rx] = (uint16_t)m;
    }
    __syncthreads();
    // non-max suppression; each warp covers 32 consecutive pixels of a row -> one plane word
    const int lane = tid & 31, warp = tid >> 5;
    for (int job = warp; job < FTH * (FTW / 32); job += 8) {
        int ry = job / (FTW / 32), wx = job - ry * (FTW / 32);
        int rx = wx * 32 + lane;
        int y = y0 + ry, x = x0 + rx;
        bool keep = false, is_strong = false;
        if (y < h && x < w) {
            int m = mg[ry + 1][rx + 1];
            if (m > low) {
                short2 d = dxy[ry][rx];
                int dx = d.x, dy = d.y;
                int ax = abs(dx), ay = abs(dy) << 15;
                int tg22x = ax * 13573;
                int tg67x = tg22x + (ax << 16);
                if (ay < tg22x) {
                    keep = (m > mg[ry + 1][rx]) && (m >= mg[ry + 1][rx + 2]);
                } else if (ay > tg67x) {
                    keep = (m > mg[ry][rx + 1]) && (m >= mg[ry + 2][rx + 1]);
                } else {
                    int sgn = ((dx ^ dy) < 0) ? -1 : 1;
                    keep = (m > mg[ry][rx + 1 - sgn]) && (m > mg[ry + 2][rx + 1 + sgn]);
                }
                is_strong = keep && (m > high);
            }
        }
        uint32_t bw = __ballot_sync(0xffffffffu, keep);
        uint32_t bs = __ballot_sync(0xffffffffu, is_strong);
        if (lane == 0 && y < h && (x0 + wx * 32) < w) {
            size_t o = (size_t)y * wpr + ((x0 >> 5) + wx);
            pw[o] = bw;
            ps[o] = bs;
        }
    }
}

// ------------------------------------------------------------- hysteresis ---
constexpr int HTHREADS = 256;        // threads per CTA
constexpr int HROWS_MAX = 256;       // rows per strip (halved until the strip fits in shared memory)
constexpr int HROWS_MIN = 8;

// shared-memory pitch (words) of a strip row: odd, to spread rows over the banks
__host__ __device__ static inline int strip_pitch(int wpr) { return wpr | 1; }

// A strip in shared memory:
//   se   : edge rows [rows + 2][sp]  (row 0 / rows+1 = halo rows of the neighbouring strips)
//   lpos : candidate words (weak bits that are not edges yet) as (row << 10 | word column)
//   lw   : their weak words
// Only candidate words can ever change, and on real images they are a small fraction of the
// strip, so every iteration touches just that list (one thread per candidate word, 32 pixels
// filled per step by fill_word).  Updates are monotone (bits are only set), so unsynchronised
// neighbour reads are benign; the loop ends after an iteration in which nothing changed.
struct Strip {
    uint32_t* se;
    uint32_t* lpos;
    uint32_t* lw;
};

__device__ __forceinline__ Strip carve_strip(uint32_t* sm, int hrows, int wpr, int sp) {
    Strip s;
    s.se = sm;
    s.lpos = sm + (hrows + 2) * sp;
    s.lw = s.lpos + hrows * wpr;
    return s;
}

__device__ __forceinline__ bool grow_word(uint32_t wv, uint32_t* se, int r, int c, int wpr, int sp) {
    uint32_t* mid = se + (r + 1) * sp;
    const uint32_t e = mid[c];
    if ((wv & ~e) == 0) return false;
    const uint32_t* up = mid - sp;
    const uint32_t* dn = mid + sp;
    uint32_t v = up[c] | e | dn[c];
    uint32_t spread = v | (v << 1) | (v >> 1);
    if (c > 0) spread |= (up[c - 1] | mid[c - 1] | dn[c - 1]) >> 31;
    if (c + 1 < wpr) spread |= (up[c + 1] | mid[c + 1] | dn[c + 1]) << 31;
    const uint32_t ne = e | (wv & spread);
    if (ne == e) return false;
    mid[c] = fill_word(ne, wv);
    return true;
}

// Load the edge rows (+ halos) of a strip and build its candidate list.  Returns the list length.
__device__ int load_strip(const uint32_t* __restrict__ weak, const uint32_t* edges, int h, int wpr, int sp, int r0,
                          int rows, const Strip& st, int* s_n) {
    if (threadIdx.x == 0) *s_n = 0;
    for (int i = threadIdx.x; i < (rows + 2) * wpr; i += blockDim.x) {
        int r = i / wpr, c = i - r * wpr;
        int rr = r0 - 1 + r;
        uint32_t v = 0;
        if (rr >= 0 && rr < h) v = __ldcg(edges + (size_t)rr * wpr + c);
        st.se[r * sp + c] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < rows * wpr; i += blockDim.x) {
        int r = i / wpr, c = i - r * wpr;
        const uint32_t wv = weak[(size_t)(r0 + r) * wpr + c];
        if (wv & ~st.se[(r + 1) * sp + c]) {
            const int k = atomicAdd(s_n, 1);
            st.lpos[k] = ((uint32_t)r << 10) | (uint32_t)c;
            st.lw[k] = wv;
        }
    }
    __syncthreads();
    return *s_n;
}

__device__ bool converge_strip(const Strip& st, int n, int wpr, int sp) {
    bool any = false;
    for (;;) {
        bool changed = false;
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const uint32_t pos = st.lpos[j];
            changed |= grow_word(st.lw[j], st.se, (int)(pos >> 10), (int)(pos & 1023u), wpr, sp);
        }
        if (!__syncthreads_or(changed)) break;
        any = true;
    }
    return any;
}

// Write the changed candidate words back; reports whether the strip's first / last row changed
// (-> the neighbouring strips' halos are stale and they must be revisited).
__device__ void store_strip(uint32_t* pe, const Strip& st, int n, int wpr, int sp, int r0, int rows, bool& top, bool& bot) {
    bool t = false, b = false;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const uint32_t pos = st.lpos[j];
        const int r = (int)(pos >> 10), c = (int)(pos & 1023u);
        const uint32_t v = st.se[(r + 1) * sp + c];
        const size_t o = (size_t)(r0 + r) * wpr + c;
        if (v != pe[o]) {
            pe[o] = v;
            t |= r == 0;
            b |= r == rows - 1;
        }
    }
    top = __syncthreads_or(t);
    bot = __syncthreads_or(b);
}

// Phase 1: strips in parallel.  `edges` enters holding the strong plane and is updated in
// place.  With flags_in only the flagged strips run.  flags[img][strip] = 1 tells that a
// neighbouring strip must be revisited.
__global__ void __launch_bounds__(HTHREADS) k_hyst_strips(const uint32_t* __restrict__ weak, uint32_t* edges, int h,
                                                          int wpr, int hrows, int nstrips, const uint32_t* flags_in,
                                                          uint32_t* flags) {
    extern __shared__ uint32_t sm[];
    __shared__ int s_n;
    const int img = blockIdx.y, strip = blockIdx.x;
    if (flags_in && flags_in[(size_t)img * nstrips + strip] == 0) return;  // later rounds: flagged strips only
    const size_t plane = (size_t)h * wpr;
    const uint32_t* pw = weak + img * plane;
    uint32_t* pe = edges + img * plane;
    const int sp = strip_pitch(wpr);
    const int r0 = strip * hrows, rows = min(hrows, h - r0);
    const Strip st = carve_strip(sm, hrows, wpr, sp);
    const int n = load_strip(pw, pe, h, wpr, sp, r0, rows, st, &s_n);
    if (n == 0 || !converge_strip(st, n, wpr, sp)) return;
    bool top, bot;
    store_strip(pe, st, n, wpr, sp, r0, rows, top, bot);
    if (threadIdx.x == 0) {
        uint32_t* f = flags + (size_t)img * nstrips;
        if (top && strip > 0) f[strip - 1] = 1;
        if (bot && strip + 1 < nstrips) f[strip + 1] = 1;
    }
}

// Phase 2: one CTA per image finishes the cross-strip propagation: visit flagged strips
// (down sweep, then up sweep) until no flag is left.  Usually there is nothing to do.
__global__ void __launch_bounds__(HTHREADS) k_hyst_finish(const uint32_t* __restrict__ weak, uint32_t* edges, int h,
                                                          int wpr, int hrows, int nstrips, uint32_t* flags) {
    extern __shared__ uint32_t sm[];
    __shared__ int s_n;
    const int img = blockIdx.x;
    const size_t plane = (size_t)h * wpr;
    const uint32_t* pw = weak + img * plane;
    uint32_t* pe = edges + img * plane;
    volatile uint32_t* f = flags + (size_t)img * nstrips;
    const int sp = strip_pitch(wpr);
    const Strip st = carve_strip(sm, hrows, wpr, sp);
    for (;;) {
        bool mine = false;
        for (int i = threadIdx.x; i < nstrips; i += blockDim.x) mine |= f[i] != 0;
        if (!__syncthreads_or(mine)) break;
        for (int pass = 0; pass < 2; ++pass) {
            for (int k = 0; k < nstrips; ++k) {
                int strip = pass == 0 ? k : nstrips - 1 - k;
                __syncthreads();
                if (f[strip] == 0) continue;  // uniform: every thread reads it after the barrier
                __syncthreads();
                if (threadIdx.x == 0) f[strip] = 0;
                const int r0 = strip * hrows, rows = min(hrows, h - r0);
                const int n = load_strip(pw, pe, h, wpr, sp, r0, rows, st, &s_n);
                if (n == 0 || !converge_strip(st, n, wpr, sp)) continue;
                bool top, bot;
                store_strip(pe, st, n, wpr, sp, r0, rows, top, bot);
                if (threadIdx.x == 0) {
                    if (top && strip > 0) f[strip - 1] = 1;
                    if (bot && strip + 1 < nstrips) f[strip + 1] = 1;
                }
            }
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------- expand ---
// bit plane -> u8 mask (0 / 255); DILATE: 3x3 max (out-of-image neighbours ignored).
// One thread per 16 output pixels (half a plane word).
template <bool DILATE>
__global__ void __launch_bounds__(256) k_plane_to_mask(const uint32_t* __restrict__ plane, int h, int w, int wpr,
                                                       uint8_t* __restrict__ mask, int aligned) {
    const int img = blockIdx.z;
    const uint32_t* p = plane + (size_t)img * h * wpr;
    uint8_t* m = mask + (size_t)img * h * w;
    const int halves = 2 * wpr;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (idx >= halves) return;
    int c = idx >> 1, half = idx & 1;
    uint32_t v;
    if (DILATE) {
        uint32_t ctr = 0, lft = 0, rgt = 0;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            int yy = y + dy;
            if (yy < 0 || yy >= h) continue;
            const uint32_t* row = p + (size_t)yy * wpr;
            ctr |= row[c];
            if (c > 0) lft |= row[c - 1];
            if (c + 1 < wpr) rgt |= row[c + 1];
        }
        v = ctr | (ctr << 1) | (ctr >> 1) | (lft >> 31) | (rgt << 31);
    } else {
        v = p[(size_t)y * wpr + c];
    }
    uint32_t bits = (v >> (16 * half)) & 0xffffu;
    int x = c * 32 + half * 16;
    if (x >= w) return;
    // 4 bits -> 4 bytes of 0x00 / 0xff
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t nib = (bits >> (4 * k)) & 0xfu;
        uint32_t spread = (nib * 0x00204081u) & 0x01010101u;  // bit i -> byte i
        o[k] = spread * 255u;
    }
    uint8_t* dst = m + (size_t)y * w + x;
    if (aligned && x + 16 <= w) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
        for (int k = 0; k < 16 && x + k < w; ++k) dst[k] = (uint8_t)(o[k >> 2] >> (8 * (k & 3)));
    }
}

// u8 mask (non-zero = set) -> bit plane; one warp per plane word
__global__ void __launch_bounds__(256) k_mask_to_plane(const uint8_t* __restrict__ mask,
                                                       const uint8_t* __restrict__ and_mask, int h, int w, int wpr,
                                                       uint32_t* __restrict__ plane) {
    const int img = blockIdx.z, y = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= wpr) return;
    const int x = c * 32 + lane;
    const size_t o = ((size_t)img * h + y) * w + x;
    const bool on = x < w && mask[o] != 0 && (and_mask == nullptr || and_mask[o] != 0);
    const uint32_t bits = __ballot_sync(0xffffffffu, on);
    if (lane == 0) plane[((size_t)img * h + y) * wpr + c] = bits;
}

// generic 3x3 max on u8 (standalone cv2.dilate with a 3x3 ones kernel)
__global__ void __launch_bounds__(256) k_dilate3_u8(const uint8_t* __restrict__ src, int h, int w,
                                                    uint8_t* __restrict__ dst) {
    const int img = blockIdx.z;
    const uint8_t* s = src + (size_t)img * h * w;
    uint8_t* d = dst + (size_t)img * h * w;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    int m = 0;
    for (int dy = -1; dy <= 1; ++dy) {
        int yy = y + dy;
        if (yy < 0 || yy >= h) continue;
        for (int dx = -1; dx <= 1; ++dx) {
            int xx = x + dx;
            if (xx < 0 || xx >= w) continue;
            m = max(m, (int)s[(size_t)yy * w + xx]);
        }
    }
    d[(size_t)y * w + x] = (uint8_t)m;
}

}  // namespace

int launch_canny_front(llfe_ctx* ctx, const uint8_t* gray, int n, int h, int w, int low, int high, uint32_t* weak,
                       uint32_t* strong) {
    if (n == 0 || h == 0 || w == 0) return LLFE_OK;
    dim3 grid(ceil_div(w, FTW), ceil_div(h, FTH), n);
    LLFE_KERNEL(ctx, "k_canny_front");
    k_canny_front<<<grid, 256, 0, ctx->stream>>>(gray, h, w, low, high, weak, strong);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

size_t hysteresis_flag_words(int n, int h) { return 2 * (size_t)n * ceil_div(h, HROWS_MIN); }

static size_t strip_smem(int hrows, int wpr) {
    return ((size_t)(hrows + 2) * strip_pitch(wpr) + 2 * (size_t)hrows * wpr) * sizeof(uint32_t);
}

// `edges` holds the strong plane on entry and the final edge plane on exit.
int launch_hysteresis(llfe_ctx* ctx, const uint32_t* weak, uint32_t* edges, int n, int h, int w, uint32_t* flags) {
    if (n == 0 || h == 0 || w == 0) return LLFE_OK;
    const int wpr = plane_wpr(w);
    if (wpr > 1023) {
        llfe_set_error("hysteresis: image width %d is beyond the 32736-pixel limit of the strip kernel", w);
        return LLFE_E_UNSUPPORTED;
    }
    const size_t budget = ctx->smem_optin - 1024;
    int hrows = HROWS_MAX;
    while (hrows > HROWS_MIN && (strip_smem(hrows, wpr) > budget / 2 || hrows >= 2 * h)) hrows >>= 1;
    const size_t smem = strip_smem(hrows, wpr);
    if (smem > budget) {
        llfe_set_error("hysteresis: image width %d needs %zu B of shared memory per strip (limit %zu)", w, smem, budget);
        return LLFE_E_UNSUPPORTED;
    }
    const int nstrips = ceil_div(h, hrows);
    if (llfe_first_use(ctx, (const void*)k_hyst_strips)) {
        LLFE_CUDA(cudaFuncSetAttribute(k_hyst_strips, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        LLFE_CUDA(cudaFuncSetAttribute(k_hyst_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    }
    // round 0 visits every strip; rounds 1..2 only strips whose halo changed; one CTA per image
    // then finishes whatever cross-strip propagation is left (usually nothing).
    const size_t fw = (size_t)n * nstrips;
    uint32_t* fa = flags;
    uint32_t* fb = flags + fw;
    LLFE_CUDA(cudaMemsetAsync(flags, 0, 2 * fw * sizeof(uint32_t), ctx->stream));
    LLFE_KERNEL(ctx, "k_hyst_strips");
    k_hyst_strips<<<dim3(nstrips, n), HTHREADS, smem, ctx->stream>>>(weak, edges, h, wpr, hrows, nstrips, nullptr, fa);
    LLFE_LAUNCHED(ctx);
    const int extra_rounds = nstrips > 1 ? 2 : 0;
    for (int round = 1; round <= extra_rounds; ++round) {
        LLFE_KERNEL(ctx, "k_hyst_strips_again");
        k_hyst_strips<<<dim3(nstrips, n), HTHREADS, smem, ctx->stream>>>(weak, edges, h, wpr, hrows, nstrips, fa, fb);
        LLFE_LAUNCHED(ctx);
        LLFE_CUDA(cudaMemsetAsync(fa, 0, fw * sizeof(uint32_t), ctx->stream));
        uint32_t* t = fa;
        fa = fb;
        fb = t;
    }
    if (nstrips > 1) {
        LLFE_KERNEL(ctx, "k_hyst_finish");
        k_hyst_finish<<<n, HTHREADS, smem, ctx->stream>>>(weak, edges, h, wpr, hrows, nstrips, fa);
        LLFE_LAUNCHED(ctx);
    }
    return LLFE_OK;
}

int launch_plane_to_mask(llfe_ctx* ctx, const uint32_t* plane, int n, int h, int w, int dilate, uint8_t* mask) {
    if (n == 0 || h == 0 || w == 0) return LLFE_OK;
    const int wpr = plane_wpr(w);
    int aligned = (w % 16 == 0) && ((uintptr_t)mask % 16 == 0);
    dim3 grid(ceil_div(2 * wpr, 256), h, n);
    LLFE_KERNEL(ctx, dilate ? "k_plane_to_mask_dilate" : "k_plane_to_mask");
    if (dilate)
        k_plane_to_mask<true><<<grid, 256, 0, ctx->stream>>>(plane, h, w, wpr, mask, aligned);
    else
        k_plane_to_mask<false><<<grid, 256, 0, ctx->stream>>>(plane, h, w, wpr, mask, aligned);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

// and_mask (optional): a pixel is set only where both maps are non-zero
int launch_mask_to_plane(llfe_ctx* ctx, const uint8_t* mask, const uint8_t* and_mask, int n, int h, int w, uint32_t* plane) {
    if (n == 0 || h == 0 || w == 0) return LLFE_OK;
    const int wpr = plane_wpr(w);
    LLFE_KERNEL(ctx, "k_mask_to_plane");
    k_mask_to_plane<<<dim3(ceil_div(wpr, 8), h, n), 256, 0, ctx->stream>>>(mask, and_mask, h, w, wpr, plane);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

int launch_dilate3_u8(llfe_ctx* ctx, const uint8_t* src, int n, int h, int w, uint8_t* dst) {
    if (n == 0 || h == 0 || w == 0) return LLFE_OK;
    dim3 grid(ceil_div(w, 256), h, n);
    LLFE_KERNEL(ctx, "k_dilate3_u8");
    k_dilate3_u8<<<grid, 256, 0, ctx->stream>>>(src, h, w, dst);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}
