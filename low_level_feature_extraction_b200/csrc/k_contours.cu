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
    int cut_shift;          // segments: cut rows are the rows y with y % (1 << cut_shift) == 0
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
    g.cut_shift = 6;
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
    int entries = 0;          // points + segment references (a reference is one entry: x = -1 - first point, y = points)
    int refs = 0;
    bool on, failed = false;
    __device__ __forceinline__ void put(int x, int y) {
        if (!on || failed) return;
        ++entries;
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

// ---- segments: long borders followed in parallel ----------------------------------------------------------------------
// The follower's state is (pixel, direction of the move into it) and its successor function is local, so ANY state is a
// valid place to start following.  A HEAD is a state on a cut row (y % 64 == 0) entered by a move with a vertical
// component, or on a cut column (x % 64 == 0) entered by a move with a horizontal one.  k_ct_segments follows every
// candidate head (foreground pixel on a cut row / column x six directions with a foreground predecessor) to the next head and records where it ends, its points, their Green sum and bounding box; the
// leader (ct_trace) then hops from head to head instead of walking.  Candidates that lie on hole borders or on no border
// at all are followed as well (nobody hops through them); a segment that is too long or finds no room for its points is
// marked invalid (end = -1) and the leader walks that stretch itself, so pruning and capacities never change the result.
// tools/debug/contour_segments_proto.py is the same algorithm in Python, checked against cv2.
// rows / columns between cuts: 1 << CtGeom::cut_shift (64; llfe_set_option("contour_cut_shift") for tests)
constexpr int CT_SEG_STEPS = 384;    // longest segment that is recorded
constexpr int CT_SEG_BUF = 8;         // points kept in registers / local memory during the first walk
constexpr int CT_SEG_MAX_STARTS = 2048;   // masks with more external components than this have short borders: no segments

struct CtSeg {
    int end;            // state id of the head this segment ends at, -1 = not recorded
    int n;              // points (direction changes) from the head (inclusive) to the end (exclusive)
    int voff;           // first point in the segment point pool
    int first, last;    // first / last point: x | y << 16
    int bb_min, bb_max; // x | y << 16
    int pad;
    long long area2;    // Green sum over consecutive points inside the segment
};
static_assert(sizeof(CtSeg) == 40, "CtSeg is read as five 8-byte words");

// State ids.  Row heads (cut row, move with a vertical component: directions 1,2,3,5,6,7) come first,
//   id = ((y / C) * 6 + code) * w + x,                      R = ceil(h / C) * 6 * w of them,
// then column heads (cut column, move with a horizontal component: directions 0,1,7,3,4,5) that are not row heads too,
//   id = R + ((x / C) * 6 + code) * h + y.
__device__ __forceinline__ int ct_ycode(int d) { return d < 4 ? d - 1 : d - 2; }       // 1,2,3,5,6,7 -> 0..5
__device__ __forceinline__ int ct_ycode_inv(int c) { return c < 3 ? c + 1 : c + 2; }
__device__ __forceinline__ int ct_xcode(int d) { return d == 7 ? 2 : d; }              // 0,1,7,3,4,5 -> 0,1,2,3,4,5
__device__ __forceinline__ int ct_xcode_inv(int c) { return c == 2 ? 7 : c; }
__device__ __forceinline__ int ct_row_states(const CtGeom& g) { return ((g.h + (1 << g.cut_shift) - 1) >> g.cut_shift) * 6 * g.w; }
__device__ __forceinline__ int ct_col_states(const CtGeom& g) { return ((g.w + (1 << g.cut_shift) - 1) >> g.cut_shift) * 6 * g.h; }

// id of the state "image pixel (x, y) entered by a move in direction d", -1 if it is not a head
__device__ __forceinline__ int ct_head_id(const CtGeom& g, int x, int y, int d) {
    const int m = (1 << g.cut_shift) - 1;
    if ((d & 3) != 0 && (y & m) == 0) return ((y >> g.cut_shift) * 6 + ct_ycode(d)) * g.w + x;
    if ((d & 3) != 2 && (x & m) == 0) return ct_row_states(g) + ((x >> g.cut_shift) * 6 + ct_xcode(d)) * g.h + y;
    return -1;
}
__device__ __forceinline__ void ct_head_decode(const CtGeom& g, int id, int& x, int& y, int& d) {
    const int R = ct_row_states(g);
    if (id < R) {
        x = id % g.w;
        const int t = id / g.w;
        d = ct_ycode_inv(t % 6);
        y = (t / 6) << g.cut_shift;
    } else {
        id -= R;
        y = id % g.h;
        const int t = id / g.h;
        d = ct_xcode_inv(t % 6);
        x = (t / 6) << g.cut_shift;
    }
}

// 8-neighbour byte of padded pixel (X, Y): bit k = neighbour in chain-code direction k
__device__ __forceinline__ uint32_t ct_nb8(const uint32_t* __restrict__ P, int pw, int X, int Y) {
    const int q = (X - 1) >> 5, sh = (X - 1) & 31;
    const uint32_t* r = P + (size_t)Y * pw + q;
    const uint32_t u3 = (uint32_t)(ct_ld2(r - pw) >> sh) & 7u, m3 = (uint32_t)(ct_ld2(r) >> sh) & 7u,
                   d3 = (uint32_t)(ct_ld2(r + pw) >> sh) & 7u;
    return ((m3 >> 2) & 1u) | (((u3 >> 2) & 1u) << 1) | (((u3 >> 1) & 1u) << 2) | ((u3 & 1u) << 3) | ((m3 & 1u) << 4) |
           ((d3 & 1u) << 5) | (((d3 >> 1) & 1u) << 6) | (((d3 >> 2) & 1u) << 7);
}

// one thread per state id: blockIdx.y = (cut row, direction code) for the row heads, then (cut column, direction code)
// for the column heads; blockIdx.x * 128 + threadIdx.x = the position along the cut (no division in the decode)
__global__ void __launch_bounds__(128) k_ct_segments(const uint32_t* __restrict__ plane, CtGeom g, CtSeg* __restrict__ tab,
                                                     int2* __restrict__ segpool_all, size_t segpool_stride, int segpool_cap,
                                                     int32_t* counts) {
    const int n_cr6 = ((g.h + (1 << g.cut_shift) - 1) >> g.cut_shift) * 6;
    const bool rows = (int)blockIdx.y < n_cr6;
    const int line = rows ? blockIdx.y : blockIdx.y - n_cr6;      // cut index * 6 + direction code
    const int along = blockIdx.x * 128 + threadIdx.x;
    if (along >= (rows ? g.w : g.h)) return;
    int32_t* cnt = counts + blockIdx.z * 4;
    if (cnt[0] > CT_SEG_MAX_STARTS) return;   // the leader does not look at the table either
    const int idx = rows ? line * g.w + along : ct_row_states(g) + line * g.h + along;
    CtSeg* rec = tab + (size_t)blockIdx.z * (ct_row_states(g) + ct_col_states(g)) + idx;
    int2* segpool = segpool_all + (size_t)blockIdx.z * segpool_stride;
    const uint32_t* P = plane + blockIdx.z * g.plane_words;
    const int pw = g.pw;
    const int cut = (line / 6) << g.cut_shift, code = line % 6;
    const int x0 = rows ? along : cut, y0 = rows ? cut : along;
    const int d0 = rows ? ct_ycode_inv(code) : ct_xcode_inv(code);
    const int Xh = x0 + 32, Yh = y0 + 1;
    auto bit = [&](int X, int Y) -> uint32_t { return (__ldg(P + (size_t)Y * pw + (X >> 5)) >> (X & 31)) & 1u; };
    // (a column head that is a row head as well lives under its row id)
    if (!bit(Xh, Yh) || !bit(Xh - ct_dx(d0), Yh - ct_dy(d0)) || ct_head_id(g, x0, y0, d0) != idx ||
        ct_nb8(P, pw, Xh, Yh) == 0xffu) {
        rec->end = -1;
        return;
    }
    int n = 0, end = -1, first = 0, last = 0, minx = 0x7fff, miny = 0x7fff, maxx = 0, maxy = 0;
    long long area2 = 0;
    int2 buf[CT_SEG_BUF];
    {
        int X = Xh, Y = Yh, din = d0, px = 0, py = 0;
        for (int step = 0; step < CT_SEG_STEPS; ++step) {
            const uint32_t nb = ct_nb8(P, pw, X, Y);
            const int s = (din + 4) & 7;
            const uint32_t r = ((nb | (nb << 8)) >> (s + 1)) & 0xffu;
            const int sn = (s + __ffs(r)) & 7;
            if (sn != din) {
                const int x = X - 32, y = Y - 1;
                if (n == 0) first = x | (y << 16);
                else area2 += (long long)px * y - (long long)py * x;
                if (n < CT_SEG_BUF) buf[n] = make_int2(x, y);
                px = x, py = y;
                minx = min(minx, x), maxx = max(maxx, x), miny = min(miny, y), maxy = max(maxy, y);
                ++n;
            }
            X += ct_dx(sn), Y += ct_dy(sn), din = sn;
            end = ct_head_id(g, X - 32, Y - 1, sn);
            if (end >= 0) break;
        }
        last = px | (py << 16);
    }
    int voff = 0;
    if (end >= 0 && n > 0) {
        voff = atomicAdd(reinterpret_cast<int*>(segpool), n);   // the first entry of the pool is its fill counter
        if (voff + n > segpool_cap) {
            end = -1;
        } else if (n <= CT_SEG_BUF) {
            for (int i = 0; i < n; ++i) segpool[1 + voff + i] = buf[i];
        } else {   // second walk: the points go straight to the pool
            int X = Xh, Y = Yh, din = d0, k = 0;
            while (k < n) {
                const uint32_t nb = ct_nb8(P, pw, X, Y);
                const int s = (din + 4) & 7;
                const uint32_t r = ((nb | (nb << 8)) >> (s + 1)) & 0xffu;
                const int sn = (s + __ffs(r)) & 7;
                if (sn != din) segpool[1 + voff + k++] = make_int2(X - 32, Y - 1);
                X += ct_dx(sn), Y += ct_dy(sn), din = sn;
            }
        }
    }
    CtSeg out;
    out.end = end, out.n = n, out.voff = voff, out.first = first, out.last = last;
    out.bb_min = minx | (miny << 16), out.bb_max = maxx | (maxy << 16), out.pad = 0, out.area2 = area2;
    *rec = out;
}

// follows the border that starts at (x0, y0): number of CHAIN_APPROX_SIMPLE points; area, bounding box; points -> sink
// SEG: tab is this image's segment table (k_ct_segments); the walk hops along it and takes single steps in between
template <bool SEG>
__device__ int ct_trace(const uint32_t* __restrict__ P, const uint32_t* __restrict__ T, const CtGeom& g, int x0, int y0,
                        CtSink& sink, long long& area2, int& minx, int& miny, int& maxx, int& maxy, const CtSeg* __restrict__ tab) {
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
    int first_head = -1;
    bool hop = SEG;
    const long long guard_max = 4ll * (g.h + 2) * (g.w + 2) + 16;
    for (long long guard = 0; guard < guard_max; ++guard) {
        int hid = -1;
        if (SEG && hop) hid = ct_head_id(g, X - 32, Y - 1, prev_s);
        if (SEG && hid >= 0) {
            // a head: hop along the recorded segments up to the one that closes the border (its end is the first head
            // this walk met), which -- like a segment that was not recorded -- is walked step by step
            if (first_head < 0) first_head = hid;
            bool moved = false;
            for (;;) {
                const int2* q = reinterpret_cast<const int2*>(tab + hid);
                const int2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3), q4 = __ldg(q + 4);
                const int e = q0.x;
                if (e < 0) break;
                if (e == first_head) {
                    hop = false;
                    break;
                }
                const int sn_pts = q0.y;
                if (sn_pts > 0) {
                    const int sfx = q1.y & 0xffff, sfy = q1.y >> 16, slx = q2.x & 0xffff, sly = q2.x >> 16;
                    if (n == 0) {
                        fx = sfx;
                        fy = sfy;
                    } else {
                        area2 += (long long)px * sfy - (long long)py * sfx;
                    }
                    area2 += ((long long)(unsigned int)q4.x) | ((long long)q4.y << 32);
                    px = slx;
                    py = sly;
                    minx = min(minx, q2.y & 0xffff);
                    miny = min(miny, q2.y >> 16);
                    maxx = max(maxx, q3.x & 0xffff);
                    maxy = max(maxy, q3.x >> 16);
                    n += sn_pts;
                    sink.put(-1 - q1.x, sn_pts);
                    ++sink.refs;
                }
                hid = e;
                moved = true;
            }
            if (moved) {
                int hx, hy;
                ct_head_decode(g, hid, hx, hy, prev_s);
                X = hx + 32;
                Y = hy + 1;
                s = (prev_s + 4) & 7;
                straight = 0;
                reload(16);
                nb = neighbours();
            }
        }
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
        if (!SEG && straight >= 1) {   // (no jumps in a walk that hops: a jump could pass a head unseen)
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
template <bool SEG>
__global__ void __launch_bounds__(128) k_ct_trace(const uint32_t* __restrict__ plane, const uint32_t* __restrict__ planeT,
                                                  CtGeom g, CtHeader* hdr, int max_contours, long long min_area2, int2* points,
                                                  int max_points, int2* pool, int pool_blocks, int32_t* counts,
                                                  const CtSeg* __restrict__ segtab, int seg_states, const int2* __restrict__ segpool,
                                                  size_t segpool_stride) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    int32_t* cnt = counts + blockIdx.z * 4;
    const int n_starts = min(cnt[0], max_contours);
    if (i >= n_starts) return;
    int n = 0, ne = 0, nrefs = 0, first = -1, off = -1;
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
        if (SEG && cnt[0] <= CT_SEG_MAX_STARTS)
            n = ct_trace<true>(P, T, g, x0, y0, sink, area2, minx, miny, maxx, maxy, segtab + (size_t)blockIdx.z * seg_states);
        else
            n = ct_trace<false>(P, T, g, x0, y0, sink, area2, minx, miny, maxx, maxy, nullptr);
        ne = sink.entries;
        nrefs = sink.refs;
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
    ne = __shfl_sync(FULL, ne, 0);
    off = __shfl_sync(FULL, off, 0);
    const int2* pl = pool + (size_t)blockIdx.z * pool_blocks * 64;
    const int2* sp = segpool ? segpool + (size_t)blockIdx.z * segpool_stride + 1 : nullptr;   // entry 0 is the fill counter
    int2* out = points + (size_t)blockIdx.z * max_points + off;
    int blk = first, opos = 0;
    n = __shfl_sync(FULL, n, 0);
    nrefs = __shfl_sync(FULL, nrefs, 0);
    if (!SEG || nrefs == 0) {   // points only
        for (int done = 0; done < n; done += CT_BLOCK_PTS) {
            const int2* src = pl + (size_t)blk * 64;
            const int m = min(CT_BLOCK_PTS, n - done);
            for (int t = lane; t < m; t += 32) out[done + t] = src[t];
            blk = src[CT_BLOCK_PTS].x;
        }
        return;
    }
    // entries are points or references to a segment's points (x < 0): positions by a warp scan of the entry sizes
    for (int done = 0; done < ne; done += CT_BLOCK_PTS) {
        const int2* src = pl + (size_t)blk * 64;
        const int m = min(CT_BLOCK_PTS, ne - done);
        for (int t0 = 0; t0 < m; t0 += 32) {
            const int t = t0 + lane;
            const int2 e = t < m ? src[t] : make_int2(0, 0);
            const int size = t < m ? (e.x < 0 ? e.y : 1) : 0;
            int inc = size;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += v;
            }
            const int at = opos + inc - size;
            if (t < m && e.x >= 0) out[at] = e;
            unsigned refs = __ballot_sync(FULL, t < m && e.x < 0);
            while (refs) {
                const int l = __ffs(refs) - 1;
                refs &= refs - 1;
                const int v0 = -1 - __shfl_sync(FULL, e.x, l), cnt_l = __shfl_sync(FULL, e.y, l), at_l = __shfl_sync(FULL, at, l);
                for (int k = lane; k < cnt_l; k += 32) out[at_l + k] = sp[v0 + k];
            }
            opos += __shfl_sync(FULL, inc, 31);
        }
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
constexpr int CT_SUB = 128;         // images per union-find sub-pass
constexpr int CT_PASS = 256;        // images per pass

// segment table + segment point pool (per image), only for calls on a few images: a batch hides the latency of a long
// border behind the other images' borders
constexpr int CT_SEG_IMAGES = 2;
static int ct_seg_states(int h, int w, int cut_shift) { return ceil_div(h, 1 << cut_shift) * w * 6 + ceil_div(w, 1 << cut_shift) * h * 6; }
static size_t ct_segpool_points(int h, int w) {
    const size_t p = (size_t)h * w / 2;
    return (p < 65536 ? 65536 : p) + 1;   // + the fill counter
}
static bool ct_use_segments(const llfe_ctx* ctx, int n, int h, int w) {
    if (!(ctx->opt_contour_segments == 2 || (ctx->opt_contour_segments == 1 && n <= CT_SEG_IMAGES))) return false;
    if (h >= 32768 || w >= 32768) return false;   // points are packed as x | y << 16
    // table + pool of one image: 24 MB at 1080p, 95 MB at 4K; beyond 512 MB (16 k x 16 k) the single walker does it
    const double per = ((double)ceil_div(h, 1 << ctx->opt_contour_cut_shift) * w + (double)ceil_div(w, 1 << ctx->opt_contour_cut_shift) * h) * 6 *
                           sizeof(CtSeg) + (double)ct_segpool_points(h, w) * 8;
    return per <= 512.0 * 1024 * 1024;
}

static size_t contours_ws_bytes(int n, int sub, int h, int w, int max_contours, int max_points, bool segments, int cut_shift) {
    const CtGeom g = ct_geom(h, w);
    return (segments ? WsCarver::need((size_t)n * ct_seg_states(h, w, cut_shift) * sizeof(CtSeg)) + WsCarver::need(n * ct_segpool_points(h, w) * 8) : 0) +
           WsCarver::need(n * (g.plane_words + g.planeT_words) * 4) +
           WsCarver::need((size_t)n * ct_pool_blocks(max_contours, max_points) * 512) + 2 * WsCarver::need(sub * g.sin_words * 4) +
           WsCarver::need(sub * g.label_words * 4);
}

static int contours_pass(llfe_ctx* ctx, const uint8_t* d_mask, int n, int sub, int h, int w, int64_t min_area2,
                         int32_t* d_headers, int max_contours, int32_t* d_points, int max_points, int32_t* d_counts, void* ws,
                         bool segments) {
    CtGeom g = ct_geom(h, w);
    g.cut_shift = ctx->opt_contour_cut_shift;
    WsCarver carve(ws);
    const int seg_states = ct_seg_states(h, w, g.cut_shift);
    const size_t segpool_stride = ct_segpool_points(h, w);
    CtSeg* segtab = segments ? carve.take<CtSeg>((size_t)n * seg_states) : nullptr;
    int2* segpool = segments ? carve.take<int2>(n * segpool_stride) : nullptr;
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
    if (segments) {
        LLFE_CUDA(cudaMemset2DAsync(segpool, segpool_stride * sizeof(int2), 0, sizeof(int2), n, ctx->stream));   // the fill counters
        LLFE_KERNEL(ctx, "k_ct_segments");
        const int cut = 1 << g.cut_shift;
        k_ct_segments<<<dim3(ceil_div(w > h ? w : h, 128), (ceil_div(h, cut) + ceil_div(w, cut)) * 6, n), 128, 0, ctx->stream>>>(plane, g, segtab, segpool,
                                                                                      segpool_stride, (int)segpool_stride - 1, d_counts);
        LLFE_LAUNCHED(ctx);
    }
    LLFE_KERNEL(ctx, "k_ct_trace");
    if (segments)
        k_ct_trace<true><<<dim3(ceil_div(max_contours, 4), 1, n), 128, 0, ctx->stream>>>(
            plane, planeT, g, (CtHeader*)d_headers, max_contours, (long long)min_area2, (int2*)d_points, max_points, pool, pool_blocks,
            d_counts, segtab, seg_states, segpool, segpool_stride);
    else
        k_ct_trace<false><<<dim3(ceil_div(max_contours, 4), 1, n), 128, 0, ctx->stream>>>(
            plane, planeT, g, (CtHeader*)d_headers, max_contours, (long long)min_area2, (int2*)d_points, max_points, pool, pool_blocks,
            d_counts, nullptr, 0, nullptr, 0);
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
    sub = sub < 1 ? 1 : sub > 256 ? 256 : sub;
    int pass = n < CT_PASS ? n : CT_PASS;
    const bool segments = ct_use_segments(ctx, n, h, w);
    if (segments) {   // segment tables + point pools of a pass: at most 2 GB
        const size_t per = (size_t)ct_seg_states(h, w, ctx->opt_contour_cut_shift) * sizeof(CtSeg) + ct_segpool_points(h, w) * 8;
        const size_t fit = ((size_t)2 << 30) / per;
        if ((size_t)pass > fit) pass = fit < 1 ? 1 : (int)fit;
    }
    if (sub > pass) sub = pass;
    void* ws;
    LLFE_TRY(llfe_workspace(ctx, contours_ws_bytes(pass, sub, h, w, max_contours, max_points, segments, ctx->opt_contour_cut_shift), &ws));
    for (int i0 = 0; i0 < n; i0 += pass) {
        const int m = n - i0 < pass ? n - i0 : pass;
        LLFE_TRY(contours_pass(ctx, d_mask + (size_t)i0 * h * w, m, sub, h, w, min_area2, d_headers + (size_t)i0 * max_contours * 10,
                               max_contours, d_points ? d_points + (size_t)i0 * max_points * 2 : nullptr, max_points,
                               d_counts + (size_t)i0 * 4, ws, segments));
    }
    return LLFE_OK;
}
