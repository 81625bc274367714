// cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) on the device (ShapeAnalyzer, shape_analyzer pyc L140 /
// L76; FontDetector.detect_text_regions, font_detector.py:51-55): the contours of a mask that is already
// in HBM, so that a few hundred polygon vertices -- not the 2 MB mask -- go back to the host, and the host no longer
// scans the mask.
//
// OpenCV's result is reproduced exactly (point sequences; the host restores cv2's contour order from the start
// pixels), from three ingredients:
//   1. which components are "external".  cv2's raster scan accepts an outer border when the last border pixel it
//      passed on the row does not carry a positive mark; on every mask probed (oracle/contours.py restates the scan,
//      tests compare both with cv2) that is: the 8-connected foreground component whose raster-first pixel has the
//      OUTER background -- the 4-connected background region that reaches the image frame -- on its left.
//   2. where each border starts: the raster-first pixel of the component.
//      Both come from ONE union-find over horizontal RUNS (foreground runs joined 8-connected, background runs
//      4-connected, frame-touching background runs joined to a virtual FRAME node).  A run's node is the linear index
//      of its first pixel, the smaller index always becomes the parent, so a component's root IS its raster-first
//      pixel.  Work is proportional to the number of runs (bit planes, 32 pixels per word), not to the pixels, and it
//      does not depend on the shape (spirals cost what rectangles cost).
//   3. the border itself: Suzuki / Abe border following with OpenCV's neighbour order (clockwise search for the last
//      border pixel from the west neighbour, then counter-clockwise steps), a point whenever the step direction
//      changes (CHAIN_APPROX_SIMPLE).  One warp slot per contour; the follower keeps a 3 x 64 pixel window of the bit
//      plane in registers and jumps along axis-aligned straight edges (up to ~60 pixels per iteration: horizontal
//      edges inside the window, vertical edges through a transposed copy of the plane), which is what UI masks are
//      made of.  The polygon's doubled area (Green's formula, exact integers == 2 * cv2.contourArea) and bounding box
//      are accumulated on the way; the points go through chained scratch blocks (one pass: the point count is only
//      known at the end) and only contours at or above the reference's threshold (`if area < 100: continue`) are
//      copied, by the whole warp, into the point array the caller gets.
#include "llfe_common.cuh"
#include "llfe_device.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

struct CtGeom {
    int h, w, wpr, pw;      // row-major plane: (h + 2) rows of pw = wpr + 2 words; pixel (x, y) = bit (x & 31) of word
                            // (x >> 5) + 1 of row y + 1 (one zero word / row of padding on every side)
    int hpr, pwT;           // transposed plane: (w + 2) rows of pwT = hpr + 2 words; pixel (x, y) = bit (y & 31) of
                            // word (y >> 5) + 1 of row x + 1
    size_t plane_words, planeT_words, sin_words, label_words;   // per image
};

__host__ __device__ inline CtGeom ct_geom(int h, int w) {
    CtGeom g;
    g.h = h;
    g.w = w;
    g.wpr = (w + 31) / 32;
    g.pw = g.wpr + 2;
    g.hpr = (h + 31) / 32;
    g.pwT = g.hpr + 2;
    g.plane_words = (size_t)(h + 2) * g.pw;
    g.planeT_words = (size_t)(w + 2) * g.pwT;
    g.sin_words = (size_t)h * g.wpr;
    g.label_words = (size_t)h * w + 1;
    return g;
}

struct CtHeader {
    int32_t start;      // linear index y * w + x of the first border pixel (raster-first pixel of the component)
    int32_t npts;       // CHAIN_APPROX_SIMPLE points
    int32_t offset;     // first point in the image's point array, -1 if the points were not written (area / capacity)
    int32_t minx, miny, maxx, maxy;
    int32_t pad;
    long long area2;    // 2 * signed polygon area (Green), |area2| / 2 == cv2.contourArea
};

// ---- union-find: the larger root is linked under the smaller one ---------------------------------------------------
__device__ __forceinline__ int uf_find(int* L, int a) {
    int p = L[a];
    while (p != a) {
        const int g = L[p];
        if (g != p) atomicMin(&L[a], g);   // path halving; parents only ever move to a smaller member of the same set
        a = p;
        p = g;
    }
    return a;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    for (;;) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(&L[a], b);   // a > b: a's parent becomes b unless someone linked a meanwhile
        if (old == a) return;
        a = old;                               // a was linked elsewhere in the meantime: join that set with b's
    }
}

__device__ __forceinline__ uint32_t valid_bits(int j, int wpr, int w) {
    return (j == wpr - 1 && (w & 31)) ? (1u << (w & 31)) - 1u : FULL;
}

// first pixel (x) of the run of `wv` that contains bit b of word j; sin_j = the answer for a run that enters the word
// from the left (see k_ct_rows)
__device__ __forceinline__ int run_start(uint32_t wv, const int* sin_row, int j, int b) {
    const uint32_t z = ~wv & ((1u << b) - 1u);
    return z ? 32 * j + 32 - __clz(z) : sin_row[j];
}

// ---- mask -> padded foreground bit plane -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ct_plane(const uint8_t* __restrict__ mask, CtGeom g, uint32_t* __restrict__ plane) {
    const int y = blockIdx.y, word = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (word >= g.wpr) return;
    const int x = word * 32 + lane;
    const bool on = x < g.w && mask[(size_t)blockIdx.z * g.h * g.w + (size_t)y * g.w + x] != 0;
    const uint32_t f = __ballot_sync(FULL, on);
    if (lane == 0) plane[blockIdx.z * g.plane_words + (size_t)(y + 1) * g.pw + word + 1] = f;
}

// the same for widths that are a multiple of 32 (rows 16-byte aligned): a thread per word, 32 mask bytes as two 128-bit
// loads; four bytes become four bits with one compare, one mask and one multiply (the partial products land on distinct bits)
__device__ __forceinline__ uint32_t ct_nibble(uint32_t v) { return (((__vcmpne4(v, 0u) & 0x01010101u) * 0x01020408u) >> 24) & 0xfu; }

__global__ void __launch_bounds__(256) k_ct_plane32(const uint8_t* __restrict__ mask, CtGeom g, uint32_t* __restrict__ plane) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= g.h * g.wpr) return;
    const int y = idx / g.wpr, word = idx - y * g.wpr;
    const uint4* p = reinterpret_cast<const uint4*>(mask + (size_t)blockIdx.z * g.h * g.w + (size_t)y * g.w + word * 32);
    const uint4 a = ld_stream(p), b = ld_stream(p + 1);
    const uint32_t f = ct_nibble(a.x) | (ct_nibble(a.y) << 4) | (ct_nibble(a.z) << 8) | (ct_nibble(a.w) << 12) |
                       (ct_nibble(b.x) << 16) | (ct_nibble(b.y) << 20) | (ct_nibble(b.z) << 24) | (ct_nibble(b.w) << 28);
    plane[blockIdx.z * g.plane_words + (size_t)(y + 1) * g.pw + word + 1] = f;
}

// 32 x 32 bit blocks of the plane, transposed (one warp per block)
__global__ void __launch_bounds__(32) k_ct_transpose(const uint32_t* __restrict__ plane, CtGeom g, uint32_t* __restrict__ planeT) {
    const int bx = blockIdx.x, by = blockIdx.y, lane = threadIdx.x;
    const int y = by * 32 + lane;
    const uint32_t wv = y < g.h ? plane[blockIdx.z * g.plane_words + (size_t)(y + 1) * g.pw + bx + 1] : 0u;
    uint32_t mine = 0;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const uint32_t t = __ballot_sync(FULL, (wv >> c) & 1u);
        if (lane == c) mine = t;
    }
    const int x = bx * 32 + lane;
    if (x < g.w) planeT[blockIdx.z * g.planeT_words + (size_t)(x + 1) * g.pwT + by + 1] = mine;
}

// One warp per row: for both planes (foreground, background) the start of the run that enters each word from the
// left (`sin`), and the union-find nodes of the runs (labels[start] = start).
__global__ void __launch_bounds__(256) k_ct_rows(const uint32_t* __restrict__ plane, CtGeom g, int* __restrict__ sinF,
                                                 int* __restrict__ sinB, int* __restrict__ labels) {
    const int y = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (y >= g.h) return;
    const uint32_t* row = plane + blockIdx.z * g.plane_words + (size_t)(y + 1) * g.pw + 1;
    int* sf = sinF + blockIdx.z * g.sin_words + (size_t)y * g.wpr;
    int* sb = sinB + blockIdx.z * g.sin_words + (size_t)y * g.wpr;
    int* lab = labels + blockIdx.z * g.label_words;
    const int base = y * g.w;
    if (y == 0 && lane == 0) lab[(size_t)g.h * g.w] = g.h * g.w;   // the FRAME node
    int carry[2] = {0, 0};
    uint32_t prev_msb[2] = {0, 0};
    for (int c = 0; c < g.wpr; c += 32) {
        const int j = c + lane;
        const bool in = j < g.wpr;
        const uint32_t f = in ? row[j] : 0u;
        const uint32_t wvs[2] = {f, in ? ~f & valid_bits(j, g.wpr, g.w) : 0u};
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
            const uint32_t wv = wvs[pl];
            const uint32_t upw = __shfl_up_sync(FULL, wv, 1);
            const uint32_t pm = lane ? upw >> 31 : prev_msb[pl];
            const int value = 32 * j + 32 - __clz(~wv);
            const unsigned nonprop = __ballot_sync(FULL, wv != FULL);
            const unsigned below = nonprop & ((1u << lane) - 1u);
            const int src = below ? 31 - __clz(below) : -1;
            const int got = __shfl_sync(FULL, value, src < 0 ? 0 : src);
            const int vin = src < 0 ? carry[pl] : got;
            if (in) (pl ? sb : sf)[j] = vin;
            uint32_t st = wv & ~((wv << 1) | pm);
            while (st) {
                const int b = __ffs(st) - 1;
                st &= st - 1;
                lab[base + 32 * j + b] = base + 32 * j + b;
            }
            const int vout = wv == FULL ? vin : value;
            carry[pl] = __shfl_sync(FULL, vout, 31);
            prev_msb[pl] = __shfl_sync(FULL, wv, 31) >> 31;
        }
    }
}

// One thread per (row, word): join the runs of this row with the runs of the row above (foreground 8-connected,
// background 4-connected), and the background runs on the frame with the FRAME node.  A maximal run of
// (row above, shifted by d) & (this row) lies inside exactly one pair of runs, so one union per such run is enough.
__global__ void __launch_bounds__(256) k_ct_merge(const uint32_t* __restrict__ plane, CtGeom g, const int* __restrict__ sinF,
                                                  const int* __restrict__ sinB, int* labels) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= g.h * g.wpr) return;
    const int y = idx / g.wpr, j = idx - y * g.wpr;
    const uint32_t* P = plane + blockIdx.z * g.plane_words + (size_t)(y + 1) * g.pw + j + 1;
    int* L = labels + blockIdx.z * g.label_words;
    const int* sfB = sinF + blockIdx.z * g.sin_words + (size_t)y * g.wpr;
    const int* sbB = sinB + blockIdx.z * g.sin_words + (size_t)y * g.wpr;
    const int* sfA = sfB - g.wpr;
    const int* sbA = sbB - g.wpr;
    const int rowB = y * g.w, rowA = rowB - g.w;
    const uint32_t B = P[0];
    const uint32_t vb = valid_bits(j, g.wpr, g.w);
    const uint32_t Bb = ~B & vb;
    if (y > 0) {
        const uint32_t A = P[-g.pw], Al = P[-g.pw - 1], Ar = P[-g.pw + 1];
        if (B) {
            uint32_t c0 = A & B;
            c0 &= ~(c0 << 1);
            while (c0) {
                const int b = __ffs(c0) - 1;
                c0 &= c0 - 1;
                uf_union(L, rowA + run_start(A, sfA, j, b), rowB + run_start(B, sfB, j, b));
            }
            uint32_t c1 = ((A << 1) | (Al >> 31)) & B & ~A;   // row above set at x - 1, clear at x (bits are isolated)
            while (c1) {
                const int b = __ffs(c1) - 1;
                c1 &= c1 - 1;
                const int sa = b ? run_start(A, sfA, j, b - 1) : run_start(Al, sfA, j - 1, 31);
                uf_union(L, rowA + sa, rowB + run_start(B, sfB, j, b));
            }
            uint32_t c2 = ((A >> 1) | (Ar << 31)) & B & ~A;   // row above set at x + 1, clear at x
            while (c2) {
                const int b = __ffs(c2) - 1;
                c2 &= c2 - 1;
                const int sa = b < 31 ? run_start(A, sfA, j, b + 1) : run_start(Ar, sfA, j + 1, 0);
                uf_union(L, rowA + sa, rowB + run_start(B, sfB, j, b));
            }
        }
        const uint32_t Ab = ~A & vb;
        uint32_t c = Ab & Bb;
        c &= ~(c << 1);
        while (c) {
            const int b = __ffs(c) - 1;
            c &= c - 1;
            uf_union(L, rowA + run_start(Ab, sbA, j, b), rowB + run_start(Bb, sbB, j, b));
        }
    }
    const int frame = g.h * g.w;
    if (y == 0 || y == g.h - 1) {
        uint32_t st = Bb & ~(Bb << 1);
        while (st) {
            const int b = __ffs(st) - 1;
            st &= st - 1;
            uf_union(L, rowB + run_start(Bb, sbB, j, b), frame);
        }
    } else {
        if (j == 0 && (Bb & 1u)) uf_union(L, rowB, frame);
        const int lb = (g.w - 1) & 31;
        if (j == g.wpr - 1 && ((Bb >> lb) & 1u)) uf_union(L, rowB + run_start(Bb, sbB, j, lb), frame);
    }
}

// starts of the external components: foreground roots whose left neighbour is outer background (or the frame)
__global__ void __launch_bounds__(256) k_ct_starts(const uint32_t* __restrict__ plane, CtGeom g, const int* __restrict__ sinB,
                                                   int* labels, CtHeader* hdr, int max_contours, int32_t* counts) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= g.h * g.wpr) return;
    const int y = idx / g.wpr, j = idx - y * g.wpr;
    const uint32_t* P = plane + blockIdx.z * g.plane_words + (size_t)(y + 1) * g.pw + j + 1;
    const uint32_t f = P[0];
    uint32_t st = f & ~((f << 1) | (P[-1] >> 31));
    if (!st) return;
    int* L = labels + blockIdx.z * g.label_words;
    const int* sb = sinB + blockIdx.z * g.sin_words + (size_t)y * g.wpr;
    int frame_root = -1;
    while (st) {
        const int b = __ffs(st) - 1;
        st &= st - 1;
        const int x = 32 * j + b, p = y * g.w + x;
        if (L[p] != p) continue;
        bool ext = x == 0;
        if (!ext) {
            const int jj = (x - 1) >> 5, bb = (x - 1) & 31;
            const uint32_t bgw = ~(jj == j ? f : P[-1]) & valid_bits(jj, g.wpr, g.w);
            if (frame_root < 0) frame_root = uf_find(L, g.h * g.w);
            ext = uf_find(L, y * g.w + run_start(bgw, sb, jj, bb)) == frame_root;
        }
        if (ext) {
            const int slot = atomicAdd(&counts[blockIdx.z * 4], 1);
            if (slot < max_contours) hdr[(size_t)blockIdx.z * max_contours + slot].start = p;
        }
    }
}

// ---- border following --------------------------------------------------------------------------------------------------
// chain code s: 0 = E, 1 = NE, 2 = N, 3 = NW, 4 = W, 5 = SW, 6 = S, 7 = SE (y grows downwards)
__device__ __forceinline__ int ct_dx(int s) { return (int)((0x901Au >> (2 * s)) & 3u) - 1; }
__device__ __forceinline__ int ct_dy(int s) { return (int)((0xA901u >> (2 * s)) & 3u) - 1; }
__device__ __forceinline__ uint64_t ct_ld2(const uint32_t* p) { return (uint64_t)__ldg(p) | ((uint64_t)__ldg(p + 1) << 32); }

// k more straight steps "upwards" in bit order (E along a bottom edge, S along a left edge): `side` must stay clear over
// [pos - 1, pos + k] and `line` set over [pos + 1, pos + k]
__device__ __forceinline__ int ct_skip_up(uint64_t side, uint64_t line, int pos) {
    const uint64_t zs = side >> (pos - 1);
    const int zrun = zs ? __ffsll((long long)zs) - 1 : 65 - pos;
    const uint64_t ol = ~(line >> (pos + 1));
    const int orun = __ffsll((long long)ol) - 1;
    return min(min(zrun - 2, orun), 61 - pos);
}
// the same "downwards" (W along a top edge, N along a right edge): `side` clear over [pos - k, pos + 1], `line` set over
// [pos - k, pos - 1]
__device__ __forceinline__ int ct_skip_down(uint64_t side, uint64_t line, int pos) {
    const uint64_t zs = side << (62 - pos);
    const int zrun = zs ? __clzll((long long)zs) : pos + 2;
    const uint64_t ol = ~(line << (64 - pos));
    const int orun = __clzll((long long)ol);
    return min(min(zrun - 2, orun), pos - 2);
}

// Points leave the follower through chained blocks of a scratch pool (the number of points of a border is only known
// when the walk is over): 63 points + the index of the next block per 512-byte block.
constexpr int CT_BLOCK_PTS = 63;

struct CtSink {
    int2* pool;
    int* pool_used;
    int pool_blocks;
    int first = -1, cur = -1, slot = 0;
    bool on, failed = false;
    __device__ __forceinline__ void put(int x, int y) {
        if (!on || failed) return;
        if (cur < 0 || slot == CT_BLOCK_PTS) {
            const int nb = atomicAdd(pool_used, 1);
            if (nb >= pool_blocks) {
                failed = true;
                return;
            }
            if (cur >= 0) pool[(size_t)cur * 64 + CT_BLOCK_PTS].x = nb;
            else first = nb;
            cur = nb;
            slot = 0;
        }
        pool[(size_t)cur * 64 + slot++] = make_int2(x, y);
    }
};

__device__ __forceinline__ void ct_prefetch(const uint32_t* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// follows the border that starts at (x0, y0): number of CHAIN_APPROX_SIMPLE points; area, bounding box; points -> sink
__device__ int ct_trace(const uint32_t* __restrict__ P, const uint32_t* __restrict__ T, const CtGeom& g, int x0, int y0,
                        CtSink& sink, long long& area2, int& minx, int& miny, int& maxx, int& maxy) {
    const int pw = g.pw, pwT = g.pwT;
    int n = 0, fx = 0, fy = 0, px = 0, py = 0;
    auto emit = [&](int x, int y) {
        sink.put(x, y);
        if (n == 0) {
            fx = x;
            fy = y;
        } else {
            area2 += (long long)px * y - (long long)py * x;
        }
        px = x;
        py = y;
        minx = min(minx, x);
        maxx = max(maxx, x);
        miny = min(miny, y);
        maxy = max(maxy, y);
        ++n;
    };
    int X = x0 + 32, Y = y0 + 1;   // padded coordinates
    int cq = 0;
    uint64_t up = 0, mid = 0, dn = 0;
    auto reload = [&](int bias) {
        cq = max(X - bias, 0) >> 5;
        const uint32_t* r = P + (size_t)Y * pw + cq;
        up = ct_ld2(r - pw);
        mid = ct_ld2(r);
        dn = ct_ld2(r + pw);
    };
    auto neighbours = [&]() -> uint32_t {
        const int sh = X - (cq << 5) - 1;
        const uint32_t u3 = (uint32_t)(up >> sh) & 7u, m3 = (uint32_t)(mid >> sh) & 7u, d3 = (uint32_t)(dn >> sh) & 7u;
        return ((m3 >> 2) & 1u) | (((u3 >> 2) & 1u) << 1) | (((u3 >> 1) & 1u) << 2) | ((u3 & 1u) << 3) | ((m3 & 1u) << 4) |
               ((d3 & 1u) << 5) | (((d3 >> 1) & 1u) << 6) | (((d3 >> 2) & 1u) << 7);
    };
    reload(16);
    uint32_t nb = neighbours();
    if (nb == 0) {   // isolated pixel
        emit(x0, y0);
        return n;
    }
    int s = 3;   // clockwise from the west neighbour: NW, N, NE (clear: raster-first pixel), E, SE, S, SW
    while (!((nb >> s) & 1u)) s = (s - 1) & 7;
    const int X0 = X, Y0 = Y, X1 = X + ct_dx(s), Y1 = Y + ct_dy(s);
    int prev_s = s ^ 4, straight = 0;
    const long long guard_max = 4ll * (g.h + 2) * (g.w + 2) + 16;
    for (long long guard = 0; guard < guard_max; ++guard) {
        const uint32_t r = ((nb | (nb << 8)) >> (s + 1)) & 0xffu;
        const int sn = (s + __ffs(r)) & 7;   // s + 1 + (ffs - 1)
        if (sn != prev_s) {
            emit(X - 32, Y - 1);
            straight = 0;
        } else {
            ++straight;
        }
        prev_s = sn;
        const int X4 = X + ct_dx(sn), Y4 = Y + ct_dy(sn);
        if (X4 == X0 && Y4 == Y0 && X == X1 && Y == Y1) break;
        if (Y4 > Y) {
            const uint32_t* r2 = P + (size_t)(Y4 + 1) * pw + cq;
            up = mid;
            mid = dn;
            dn = ct_ld2(r2);
            ct_prefetch(r2 + 2 * pw);   // borders are locally coherent: the rows ahead will be wanted next
        } else if (Y4 < Y) {
            const uint32_t* r2 = P + (size_t)(Y4 - 1) * pw + cq;
            dn = mid;
            mid = up;
            up = ct_ld2(r2);
            if (Y4 >= 3) ct_prefetch(r2 - 2 * pw);
        }
        X = X4;
        Y = Y4;
        s = (sn + 4) & 7;
        int pos = X - (cq << 5);
        if (pos < 2 || pos > 61) {
            reload(sn == 0 ? 2 : sn == 4 ? 30 : 16);
            pos = X - (cq << 5);
        }
        if (straight >= 1) {
            if (sn == 0) {
                const int k = ct_skip_up(dn, mid, pos);
                if (k > 0) X += k;
            } else if (sn == 4) {
                int k = ct_skip_down(up, mid, pos);
                if (Y == Y0 && X > X0) k = min(k, X - X0 - 1);   // the step into the start pixel ends the border: take it normally
                if (k > 0) X -= k;
            } else if (straight >= 2 && (sn == 6 || sn == 2)) {
                const int x = X - 32, Yt = Y + 31;
                const int cqT = max(Yt - (sn == 6 ? 2 : 30), 0) >> 5, posT = Yt - (cqT << 5);
                const uint64_t line = ct_ld2(T + (size_t)(x + 1) * pwT + cqT);
                int k;
                if (sn == 6) {
                    k = ct_skip_up(ct_ld2(T + (size_t)x * pwT + cqT), line, posT);
                } else {
                    k = ct_skip_down(ct_ld2(T + (size_t)(x + 2) * pwT + cqT), line, posT);
                    if (X == X0 && Y > Y0) k = min(k, Y - Y0 - 1);
                }
                if (k > 0) {
                    Y += sn == 6 ? k : -k;
                    reload(16);
                }
            }
        }
        nb = neighbours();
    }
    area2 += (long long)px * fy - (long long)py * fx;   // last -> first closes Green's sum
    return n;
}

// One warp per contour.  Lane 0 follows the border (the walk is sequential and latency-bound; a converged warp keeps
// other contours from serialising behind it) and leaves the points in chained scratch blocks; then the whole warp
// copies the points of a contour that passed the area threshold to its final place in the image's point array.
__global__ void __launch_bounds__(128) k_ct_trace(const uint32_t* __restrict__ plane, const uint32_t* __restrict__ planeT,
                                                  CtGeom g, CtHeader* hdr, int max_contours, long long min_area2, int2* points,
                                                  int max_points, int2* pool, int pool_blocks, int32_t* counts) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    int32_t* cnt = counts + blockIdx.z * 4;
    const int n_starts = min(cnt[0], max_contours);
    if (i >= n_starts) return;
    int n = 0, first = -1, off = -1;
    if (lane == 0) {
        CtHeader* hd = hdr + (size_t)blockIdx.z * max_contours + i;
        const uint32_t* P = plane + blockIdx.z * g.plane_words;
        const uint32_t* T = planeT + blockIdx.z * g.planeT_words;
        const int p = hd->start, y0 = p / g.w, x0 = p - y0 * g.w;
        long long area2 = 0;
        int minx = x0, miny = y0, maxx = x0, maxy = y0;
        CtSink sink;
        sink.pool = pool + (size_t)blockIdx.z * pool_blocks * 64;
        sink.pool_used = &cnt[3];
        sink.pool_blocks = pool_blocks;
        sink.on = max_points > 0;
        n = ct_trace(P, T, g, x0, y0, sink, area2, minx, miny, maxx, maxy);
        CtHeader H;
        H.start = p;
        H.npts = n;
        H.offset = -1;
        H.minx = minx;
        H.miny = miny;
        H.maxx = maxx;
        H.maxy = maxy;
        H.pad = 0;
        H.area2 = area2;
        const long long a = area2 < 0 ? -area2 : area2;
        if (a >= min_area2 && max_points > 0) {
            if (sink.failed) {
                atomicOr(&cnt[2], 1);   // scratch pool exhausted: the caller retries with larger buffers
            } else {
                off = atomicAdd(&cnt[1], n);
                if (off + n <= max_points) {
                    H.offset = off;
                    first = sink.first;
                } else {
                    atomicOr(&cnt[2], 1);   // point capacity exceeded
                    off = -1;
                }
            }
        }
        *hd = H;
    }
    __syncwarp();
    first = __shfl_sync(FULL, first, 0);
    if (first < 0) return;
    n = __shfl_sync(FULL, n, 0);
    off = __shfl_sync(FULL, off, 0);
    const int2* pl = pool + (size_t)blockIdx.z * pool_blocks * 64;
    int2* out = points + (size_t)blockIdx.z * max_points + off;
    int blk = first;
    for (int done = 0; done < n; done += CT_BLOCK_PTS) {
        const int2* src = pl + (size_t)blk * 64;
        const int m = min(CT_BLOCK_PTS, n - done);
        for (int t = lane; t < m; t += 32) out[done + t] = src[t];
        blk = src[CT_BLOCK_PTS].x;
    }
}

}  // namespace

static int ct_pool_blocks(int max_contours, int max_points) {
    return max_points > 0 ? max_contours + 2 * ceil_div(max_points, CT_BLOCK_PTS) : 0;
}

// Workspace of one pass over n images: planes + scratch point pool for all of them (every border of the pass is followed
// in ONE launch: the pass lasts as long as its longest border, so it should cover many images), union-find state (4 bytes
// per pixel: the part that must stay L2-sized) for `sub` images at a time.
constexpr int CT_SUB = 16;          // images per union-find sub-pass (1080p: 16 x 9.3 MB)
constexpr int CT_PASS = 256;        // images per pass

static size_t contours_ws_bytes(int n, int sub, int h, int w, int max_contours, int max_points) {
    const CtGeom g = ct_geom(h, w);
    return WsCarver::need(n * (g.plane_words + g.planeT_words) * 4) +
           WsCarver::need((size_t)n * ct_pool_blocks(max_contours, max_points) * 512) + 2 * WsCarver::need(sub * g.sin_words * 4) +
           WsCarver::need(sub * g.label_words * 4);
}

static int contours_pass(llfe_ctx* ctx, const uint8_t* d_mask, int n, int sub, int h, int w, int64_t min_area2,
                         int32_t* d_headers, int max_contours, int32_t* d_points, int max_points, int32_t* d_counts, void* ws) {
    const CtGeom g = ct_geom(h, w);
    WsCarver carve(ws);
    uint32_t* plane = carve.take<uint32_t>(n * (g.plane_words + g.planeT_words));
    uint32_t* planeT = plane + n * g.plane_words;
    const int pool_blocks = ct_pool_blocks(max_contours, max_points);
    int2* pool = carve.take<int2>((size_t)n * pool_blocks * 64);
    int* sinF = carve.take<int>(sub * g.sin_words);
    int* sinB = carve.take<int>(sub * g.sin_words);
    int* labels = carve.take<int>(sub * g.label_words);
    LLFE_CUDA(cudaMemsetAsync(plane, 0, n * (g.plane_words + g.planeT_words) * 4, ctx->stream));
    LLFE_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)n * 4 * sizeof(int32_t), ctx->stream));
    LLFE_KERNEL(ctx, "k_ct_plane");
    if (w % 32 == 0 && ((uintptr_t)d_mask & 15) == 0)
        k_ct_plane32<<<dim3(ceil_div(h * g.wpr, 256), 1, n), 256, 0, ctx->stream>>>(d_mask, g, plane);
    else
        k_ct_plane<<<dim3(ceil_div(g.wpr, 8), h, n), 256, 0, ctx->stream>>>(d_mask, g, plane);
    LLFE_LAUNCHED(ctx);
    LLFE_KERNEL(ctx, "k_ct_transpose");
    k_ct_transpose<<<dim3(g.wpr, g.hpr, n), 32, 0, ctx->stream>>>(plane, g, planeT);
    LLFE_LAUNCHED(ctx);
    const int words = h * g.wpr;
    for (int i0 = 0; i0 < n; i0 += sub) {
        const int m = n - i0 < sub ? n - i0 : sub;
        const uint32_t* pl = plane + (size_t)i0 * g.plane_words;
        LLFE_KERNEL(ctx, "k_ct_rows");
        k_ct_rows<<<dim3(ceil_div(h, 8), 1, m), 256, 0, ctx->stream>>>(pl, g, sinF, sinB, labels);
        LLFE_LAUNCHED(ctx);
        LLFE_KERNEL(ctx, "k_ct_merge");
        k_ct_merge<<<dim3(ceil_div(words, 256), 1, m), 256, 0, ctx->stream>>>(pl, g, sinF, sinB, labels);
        LLFE_LAUNCHED(ctx);
        LLFE_KERNEL(ctx, "k_ct_starts");
        k_ct_starts<<<dim3(ceil_div(words, 256), 1, m), 256, 0, ctx->stream>>>(
            pl, g, sinB, labels, (CtHeader*)d_headers + (size_t)i0 * max_contours, max_contours, d_counts + (size_t)i0 * 4);
        LLFE_LAUNCHED(ctx);
    }
    LLFE_KERNEL(ctx, "k_ct_trace");
    k_ct_trace<<<dim3(ceil_div(max_contours, 4), 1, n), 128, 0, ctx->stream>>>(plane, planeT, g, (CtHeader*)d_headers,
                                                                              max_contours, (long long)min_area2,
                                                                              (int2*)d_points, max_points, pool, pool_blocks,
                                                                              d_counts);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

// d_mask (n, h, w) u8 -> per image: headers (max_contours x 10 int32, see CtHeader), points (max_points x 2 int32),
// counts {external components found, points written, capacity flag, scratch blocks used}
extern "C" int llfe_contours_external(llfe_ctx* ctx, const uint8_t* d_mask, int n, int h, int w, int64_t min_area2,
                                      int32_t* d_headers, int max_contours, int32_t* d_points, int max_points,
                                      int32_t* d_counts) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_mask != nullptr && d_headers != nullptr && d_counts != nullptr && (d_points != nullptr || max_points == 0));
    LLFE_CHECK_ARG(n > 0 && h > 0 && w > 0 && (size_t)h * w < 0x7ffffff0ull && max_contours > 0 && max_points >= 0);
    static_assert(sizeof(CtHeader) == 40, "header layout is part of the ABI");
    // the union-find nodes of a sub-pass should stay in L2 (16 x 1080p), at least one image
    const size_t px = (size_t)h * w;
    int sub = (int)(((size_t)CT_SUB * 1080 * 1920) / px);
    sub = sub < 1 ? 1 : sub > 64 ? 64 : sub;
    const int pass = n < CT_PASS ? n : CT_PASS;
    if (sub > pass) sub = pass;
    void* ws;
    LLFE_TRY(llfe_workspace(ctx, contours_ws_bytes(pass, sub, h, w, max_contours, max_points), &ws));
    for (int i0 = 0; i0 < n; i0 += pass) {
        const int m = n - i0 < pass ? n - i0 : pass;
        LLFE_TRY(contours_pass(ctx, d_mask + (size_t)i0 * h * w, m, sub, h, w, min_area2, d_headers + (size_t)i0 * max_contours * 10,
                               max_contours, d_points ? d_points + (size_t)i0 * max_points * 2 : nullptr, max_points,
                               d_counts + (size_t)i0 * 4, ws));
    }
    return LLFE_OK;
}
