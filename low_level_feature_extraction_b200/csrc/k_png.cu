// PNG scanline reconstruction + conversion to the BGR image `cv2.imdecode(buf, cv2.IMREAD_COLOR)` returns
// (reference call sites: app/services/analyze/utils.py:108-109, image_processor.py:62-66, :208-211).
//
// A PNG's IDAT chunks hold ONE zlib stream; inflated it is h scanlines of [filter type byte][rowbytes filtered bytes].
// Inflate is a serial bit-level decode of a single stream and stays on the host (zlib); everything after it runs here:
//
//   k_png_unfilter   Recon(x) = Filt(x) + predictor(a = left, b = up, c = up-left), bytes `bpp` apart (PNG spec 9.2:
//                    None, Sub, Up, Average = floor((a + b) / 2), Paeth).  Every byte depends on its left, upper and
//                    upper-left neighbours, so one image is reconstructed as a skewed wavefront: thread j owns rows
//                    j, j + T, j + 2T, ... and walks each in 16-byte chunks, one chunk behind the thread that owns the row
//                    above (rows are padded to at least T chunks, so the first row of the next round never overtakes the
//                    last row of the previous one); a block barrier per step.  One CTA per image: a batch fills the
//                    SMs (148 x 1080p in 3.6 ms); ONE image is bound by the 2 056 dependent steps (3.4 ms at 1080p).
//   k_png_to_bgr     pointwise: samples of 1 / 2 / 4 / 8 / 16 bits -> 8 bits the way OpenCV configures libpng for
//                    IMREAD_COLOR (16 -> the high byte, gray 1/2/4 scaled by 255/85/17, palette looked up, alpha and tRNS
//                    dropped, gray replicated), RGB -> BGR.
#include "llfe_common.cuh"
#include "llfe_device.cuh"

namespace {

constexpr int PNG_CH = 16;               // bytes of a row per wavefront step
constexpr int PNG_TW = PNG_CH / 4 + 1;   // words per shared-memory row (odd: a lane per row reads without bank conflicts)
constexpr int PNG_CARRY = 1536;          // slots of the last-thread -> first-thread delay line

__device__ __forceinline__ int paeth(int a, int b, int c) {
    const int pa = abs(b - c), pb = abs(a - c), pc = abs(a + b - 2 * c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Position of a thread in its schedule: thread j starts row j at step j, walks it a chunk per step, idles until step
// `cpad` of the row (so that the thread above stays ahead), then continues with row j + T.
struct PngPos {
    int r, sc;
    __device__ __forceinline__ void advance(int T, int cpad) {
        if (++sc == cpad) sc = 0, r += T;
    }
    __device__ __forceinline__ bool active(int h, int chunks) const { return sc >= 0 && sc < chunks && r < h; }
};

// The latencies are kept off the per-step critical path (a block barrier closes every step, so the slowest thread
// sets the pace): global memory is touched by the whole warp, 2 row-chunks per instruction; the chunks and the filter
// byte of step t + 1 are fetched into registers while step t is reconstructed; the row above comes through shared
// memory -- from the neighbouring thread's previous step, or, for the first thread, from a delay line the last thread
// feeds (its row above was finished cpad - T + 1 steps earlier).
template <int BPP>
__global__ void __launch_bounds__(256) k_png_unfilter(uint8_t* __restrict__ stream, int h, int row0, int row1, int rowbytes,
                                                      int chunks, int cpad, int use_carry, int32_t* __restrict__ status) {
    __shared__ uint32_t tin[256 * PNG_TW];          // filtered bytes of this step, a row per thread
    __shared__ uint32_t tout[2][256 * PNG_TW];      // reconstructed bytes of this / the previous step (the row above)
    __shared__ uint4 carry[PNG_CARRY];
    const uint32_t stride = (uint32_t)rowbytes + 1u;
    uint8_t* img = stream + (size_t)blockIdx.x * h * stride;
    const int T = blockDim.x, j = threadIdx.x, lane = j & 31, wbase = j & ~31;
    const int total = T + ((row1 - row0 + T - 1) / T) * cpad;   // rows [row0, row1); rows above row0 are reconstructed already
    const int d1 = cpad - T + 2;                    // delay-line length
    const int bi = lane & 15, half = lane >> 4;
    int win[BPP], cw[BPP];   // the last BPP reconstructed bytes of this row (a) and of the row above (c)
#pragma unroll
    for (int k = 0; k < BPP; ++k) win[k] = cw[k] = 0;
    int ft = 0, ftn = 0, wr = 0;
    uint32_t vn[16];         // byte `bi` of the chunks of rows 2i + half of this warp, for the NEXT step
    PngPos cur{row0 + j, -j};
    auto fetch = [&](const PngPos& p) {
        const bool act = p.active(row1, chunks);
        const uint32_t off = act ? (uint32_t)p.r * stride + 1u + (uint32_t)p.sc * PNG_CH : 0u;
        const int nb = act ? min(PNG_CH, rowbytes - p.sc * PNG_CH) : 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t ok = __shfl_sync(0xffffffffu, off, 2 * i + half);
            const int nk = __shfl_sync(0xffffffffu, nb, 2 * i + half);
            vn[i] = bi < nk ? img[ok + bi] : 0u;
        }
        if (act && p.sc == 0) ftn = img[(uint32_t)p.r * stride];
    };
    auto stage = [&]() {
#pragma unroll
        for (int i = 0; i < 16; ++i) reinterpret_cast<uint8_t*>(tin + (wbase + 2 * i + half) * PNG_TW)[bi] = (uint8_t)vn[i];
    };
    fetch(cur);
    stage();
    __syncwarp();
    for (int t = 0; t < total; ++t) {
        PngPos nxt = cur;
        nxt.advance(T, cpad);
        const bool act = cur.active(row1, chunks);
        if (act && cur.sc == 0) ft = ftn;
        fetch(nxt);
        uint32_t* mine = tout[t & 1] + j * PNG_TW;
        const int rd = wr + 1 == d1 ? 0 : wr + 1;
        if (act) {
            const int r = cur.r, sc = cur.sc;
            const int nb = min(PNG_CH, rowbytes - sc * PNG_CH);
            if (sc == 0) {
#pragma unroll
                for (int k = 0; k < BPP; ++k) win[k] = cw[k] = 0;
            }
            int x[PNG_CH], b[PNG_CH];
#pragma unroll
            for (int w = 0; w < PNG_CH / 4; ++w) {
                const uint32_t v = tin[j * PNG_TW + w];
                x[4 * w] = v & 255, x[4 * w + 1] = (v >> 8) & 255, x[4 * w + 2] = (v >> 16) & 255, x[4 * w + 3] = v >> 24;
            }
            if (r == 0) {
#pragma unroll
                for (int i = 0; i < PNG_CH; ++i) b[i] = 0;
            } else if (j > 0 || (use_carry && r >= row0 + T)) {
                // the thread above finished this chunk of its row in the previous step; for the first thread that
                // was the last thread, d1 - 1 steps ago (in its first round: the row above is in the stream already)
                uint32_t v4[4];
                if (j > 0) {
                    const uint32_t* above = tout[(t & 1) ^ 1] + (j - 1) * PNG_TW;
#pragma unroll
                    for (int w = 0; w < 4; ++w) v4[w] = above[w];
                } else {
                    const uint4 c4 = carry[rd];
                    v4[0] = c4.x, v4[1] = c4.y, v4[2] = c4.z, v4[3] = c4.w;
                }
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const uint32_t v = v4[w];
                    b[4 * w] = v & 255, b[4 * w + 1] = (v >> 8) & 255, b[4 * w + 2] = (v >> 16) & 255, b[4 * w + 3] = v >> 24;
                }
            } else {
                // rows too long for the delay line: the first thread reads the row above back from the stream (it was
                // stored at least one step ago, because rows are padded to cpad >= T chunks)
                const uint8_t* up = img + (uint32_t)(r - 1) * stride + 1u + (uint32_t)sc * PNG_CH;
#pragma unroll
                for (int i = 0; i < PNG_CH; ++i) b[i] = i < nb ? up[i] : 0;
            }
            if (ft == 1) {
#pragma unroll
                for (int i = 0; i < PNG_CH; ++i) {
                    const int a = i < BPP ? win[i] : x[i - BPP];
                    x[i] = (x[i] + a) & 255;
                }
            } else if (ft == 2) {
#pragma unroll
                for (int i = 0; i < PNG_CH; ++i) x[i] = (x[i] + b[i]) & 255;
            } else if (ft == 3) {
#pragma unroll
                for (int i = 0; i < PNG_CH; ++i) {
                    const int a = i < BPP ? win[i] : x[i - BPP];
                    x[i] = (x[i] + ((a + b[i]) >> 1)) & 255;
                }
            } else if (ft == 4) {
#pragma unroll
                for (int i = 0; i < PNG_CH; ++i) {
                    const int a = i < BPP ? win[i] : x[i - BPP];
                    const int c = i < BPP ? cw[i] : b[i - BPP];
                    x[i] = (x[i] + paeth(a, b[i], c)) & 255;
                }
            } else if (ft != 0) {
                if (sc == 0) atomicOr(&status[blockIdx.x], 1);   // libpng: "bad adaptive filter value"
            }
            uint4 o;
            o.x = (uint32_t)x[0] | ((uint32_t)x[1] << 8) | ((uint32_t)x[2] << 16) | ((uint32_t)x[3] << 24);
            o.y = (uint32_t)x[4] | ((uint32_t)x[5] << 8) | ((uint32_t)x[6] << 16) | ((uint32_t)x[7] << 24);
            o.z = (uint32_t)x[8] | ((uint32_t)x[9] << 8) | ((uint32_t)x[10] << 16) | ((uint32_t)x[11] << 24);
            o.w = (uint32_t)x[12] | ((uint32_t)x[13] << 8) | ((uint32_t)x[14] << 16) | ((uint32_t)x[15] << 24);
            mine[0] = o.x, mine[1] = o.y, mine[2] = o.z, mine[3] = o.w;
            if (j == T - 1 && use_carry) carry[wr] = o;
            // the next chunk's left / upper-left neighbours (a full chunk always precedes another chunk)
#pragma unroll
            for (int k = 0; k < BPP; ++k) win[k] = x[PNG_CH - BPP + k], cw[k] = b[PNG_CH - BPP + k];
        }
        __syncwarp();
        // the warp stores its reconstructed chunks, two rows per instruction
        {
            const uint32_t off = act ? (uint32_t)cur.r * stride + 1u + (uint32_t)cur.sc * PNG_CH : 0u;
            const int nb = act ? min(PNG_CH, rowbytes - cur.sc * PNG_CH) : 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t ok = __shfl_sync(0xffffffffu, off, 2 * i + half);
                const int nk = __shfl_sync(0xffffffffu, nb, 2 * i + half);
                if (bi < nk) img[ok + bi] = reinterpret_cast<const uint8_t*>(tout[t & 1] + (wbase + 2 * i + half) * PNG_TW)[bi];
            }
        }
        stage();          // the prefetched chunks become the next step's input (tin rows are private to the warp)
        wr = rd;
        cur = nxt;
        __syncthreads();
    }
}

struct PngFmt {
    int h, w, rowbytes, color_type, depth;
};

__device__ __forceinline__ int png_sample(const uint8_t* row, int idx, int depth) {
    if (depth == 8) return row[idx];
    if (depth == 16) return row[2 * idx];                      // png_set_strip_16: the high byte
    const int bit = idx * depth;
    return (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
}

__global__ void __launch_bounds__(256) k_png_to_bgr(const uint8_t* __restrict__ stream, PngFmt f,
                                                    const uint8_t* __restrict__ palette, uint8_t* __restrict__ bgr) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= f.w) return;
    const size_t stride = (size_t)f.rowbytes + 1;
    const uint8_t* row = stream + ((size_t)img * f.h + y) * stride + 1;
    int r, g, b;
    switch (f.color_type) {
        case 0: {
            int v = png_sample(row, x, f.depth);
            if (f.depth < 8) v *= f.depth == 1 ? 255 : f.depth == 2 ? 85 : 17;   // png_set_expand_gray_1_2_4_to_8
            r = g = b = v;
            break;
        }
        case 2:
            r = png_sample(row, 3 * x, f.depth), g = png_sample(row, 3 * x + 1, f.depth), b = png_sample(row, 3 * x + 2, f.depth);
            break;
        case 3: {
            const uint8_t* p = palette + (size_t)img * 768 + 3 * png_sample(row, x, f.depth);
            r = p[0], g = p[1], b = p[2];
            break;
        }
        case 4:
            r = g = b = png_sample(row, 2 * x, f.depth);
            break;
        default:
            r = png_sample(row, 4 * x, f.depth), g = png_sample(row, 4 * x + 1, f.depth), b = png_sample(row, 4 * x + 2, f.depth);
            break;
    }
    uint8_t* o = bgr + (((size_t)img * f.h + y) * f.w + x) * 3;
    o[0] = (uint8_t)b, o[1] = (uint8_t)g, o[2] = (uint8_t)r;
}

// Adam7 (PNG specification 8.2): the stream holds seven reduced images, pass p covering the pixels (x0 + i * dx, y0 + j * dy)
struct PngAdam7 {
    int h, w, color_type, depth;
    unsigned int off[7];          // offset of each pass in the stream (empty passes: unused)
    int rowbytes[7];
};
__constant__ uint8_t c_adam7_pass[64] = {0, 5, 3, 5, 1, 5, 3, 5, 6, 6, 6, 6, 6, 6, 6, 6, 4, 5, 4, 5, 4, 5, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6,
                                         2, 5, 3, 5, 2, 5, 3, 5, 6, 6, 6, 6, 6, 6, 6, 6, 4, 5, 4, 5, 4, 5, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6};
__constant__ uint8_t c_adam7_geom[7][4] = {{0, 0, 3, 3}, {4, 0, 3, 3}, {0, 4, 2, 3}, {2, 0, 2, 2}, {0, 2, 1, 2}, {1, 0, 1, 1}, {0, 1, 0, 1}};   // x0, y0, log2 dx, log2 dy

__global__ void __launch_bounds__(256) k_png_adam7_to_bgr(const uint8_t* __restrict__ stream, PngAdam7 f,
                                                          const uint8_t* __restrict__ palette, uint8_t* __restrict__ bgr) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= f.w) return;
    const int p = c_adam7_pass[(y & 7) * 8 + (x & 7)];
    const int px = (x - c_adam7_geom[p][0]) >> c_adam7_geom[p][2], py = (y - c_adam7_geom[p][1]) >> c_adam7_geom[p][3];
    const uint8_t* row = stream + f.off[p] + (size_t)py * (f.rowbytes[p] + 1) + 1;
    int r, g, b;
    switch (f.color_type) {
        case 0: {
            int v = png_sample(row, px, f.depth);
            if (f.depth < 8) v *= f.depth == 1 ? 255 : f.depth == 2 ? 85 : 17;
            r = g = b = v;
            break;
        }
        case 2:
            r = png_sample(row, 3 * px, f.depth), g = png_sample(row, 3 * px + 1, f.depth), b = png_sample(row, 3 * px + 2, f.depth);
            break;
        case 3: {
            const uint8_t* q = palette + 3 * png_sample(row, px, f.depth);
            r = q[0], g = q[1], b = q[2];
            break;
        }
        case 4:
            r = g = b = png_sample(row, 2 * px, f.depth);
            break;
        default:
            r = png_sample(row, 4 * px, f.depth), g = png_sample(row, 4 * px + 1, f.depth), b = png_sample(row, 4 * px + 2, f.depth);
            break;
    }
    uint8_t* o = bgr + ((size_t)y * f.w + x) * 3;
    o[0] = (uint8_t)b, o[1] = (uint8_t)g, o[2] = (uint8_t)r;
}

}  // namespace

static int png_channels(int color_type) {
    switch (color_type) {
        case 0: return 1;
        case 2: return 3;
        case 3: return 1;
        case 4: return 2;
        case 6: return 4;
    }
    return 0;
}

extern "C" int64_t llfe_png_rowbytes(int w, int color_type, int bit_depth) {
    const int ch = png_channels(color_type);
    if (ch == 0 || w <= 0) return -1;
    const bool ok = color_type == 0   ? (bit_depth == 1 || bit_depth == 2 || bit_depth == 4 || bit_depth == 8 || bit_depth == 16)
                    : color_type == 3 ? (bit_depth == 1 || bit_depth == 2 || bit_depth == 4 || bit_depth == 8)
                                      : (bit_depth == 8 || bit_depth == 16);
    if (!ok) return -1;
    return ((int64_t)w * ch * bit_depth + 7) / 8;
}

// reconstruct rows [row0, row1) of n images in place (rows above row0 must be reconstructed already)
int launch_png_unfilter_rows(llfe_ctx* ctx, uint8_t* d_stream, int n, int h, int row0, int row1, int rowbytes, int bpp,
                             int32_t* d_status) {
    const int chunks = ceil_div(rowbytes, PNG_CH);
    const int rows = row1 - row0;
    int T = chunks >= 256 ? 256 : (chunks / 32) * 32;
    if (T < 32) T = 32;
    if (T > ((rows + 31) / 32) * 32) T = ((rows + 31) / 32) * 32;
    const int cpad = chunks > T ? chunks : T;
    const int use_carry = cpad - T + 2 <= PNG_CARRY;
    LLFE_KERNEL(ctx, "k_png_unfilter");
    switch (bpp) {
        case 1: k_png_unfilter<1><<<n, T, 0, ctx->stream>>>(d_stream, h, row0, row1, rowbytes, chunks, cpad, use_carry, d_status); break;
        case 2: k_png_unfilter<2><<<n, T, 0, ctx->stream>>>(d_stream, h, row0, row1, rowbytes, chunks, cpad, use_carry, d_status); break;
        case 3: k_png_unfilter<3><<<n, T, 0, ctx->stream>>>(d_stream, h, row0, row1, rowbytes, chunks, cpad, use_carry, d_status); break;
        case 4: k_png_unfilter<4><<<n, T, 0, ctx->stream>>>(d_stream, h, row0, row1, rowbytes, chunks, cpad, use_carry, d_status); break;
        case 6: k_png_unfilter<6><<<n, T, 0, ctx->stream>>>(d_stream, h, row0, row1, rowbytes, chunks, cpad, use_carry, d_status); break;
        case 8: k_png_unfilter<8><<<n, T, 0, ctx->stream>>>(d_stream, h, row0, row1, rowbytes, chunks, cpad, use_carry, d_status); break;
        default: llfe_set_error("llfe_png_reconstruct: unsupported pixel size %d", bpp); return LLFE_E_UNSUPPORTED;
    }
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

int launch_png_to_bgr(llfe_ctx* ctx, const uint8_t* d_stream, int n, int h, int w, int rowbytes, int color_type, int bit_depth,
                      const uint8_t* d_palette, uint8_t* d_bgr) {
    const PngFmt f{h, w, rowbytes, color_type, bit_depth};
    LLFE_KERNEL(ctx, "k_png_to_bgr");
    k_png_to_bgr<<<dim3(ceil_div(w, 256), h, n), 256, 0, ctx->stream>>>(d_stream, f, d_palette, d_bgr);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}

int png_filter_distance(int color_type, int bit_depth) { return (png_channels(color_type) * bit_depth + 7) / 8; }

extern "C" int llfe_png_reconstruct(llfe_ctx* ctx, uint8_t* d_stream, int n, int h, int w, int color_type, int bit_depth,
                                    const uint8_t* d_palette, uint8_t* d_bgr, int32_t* d_status) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_stream != nullptr && d_bgr != nullptr && d_status != nullptr);
    LLFE_CHECK_ARG(n >= 0 && n <= 65535 && h > 0 && h <= 65535 && w > 0);
    const int64_t rb = llfe_png_rowbytes(w, color_type, bit_depth);
    LLFE_CHECK_ARG(rb > 0 && rb < 0x7fffffff);
    LLFE_CHECK_ARG(color_type != 3 || d_palette != nullptr);
    if (n == 0) return LLFE_OK;
    const int rowbytes = (int)rb;
    LLFE_CHECK_ARG((uint64_t)h * ((uint64_t)rowbytes + 1) < 0xffffffffull);   // 32-bit offsets inside one image
    LLFE_CUDA(cudaMemsetAsync(d_status, 0, (size_t)n * sizeof(int32_t), ctx->stream));
    LLFE_TRY(launch_png_unfilter_rows(ctx, d_stream, n, h, 0, h, rowbytes, png_filter_distance(color_type, bit_depth), d_status));
    return launch_png_to_bgr(ctx, d_stream, n, h, w, rowbytes, color_type, bit_depth, d_palette, d_bgr);
}

// ---- Adam7-interlaced files (one image per call) ---------------------------------------------------------------------------
static const int ADAM7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};

// size of the scanline stream of a w x h PNG; interlace = IHDR's interlace method (0 or 1)
extern "C" int64_t llfe_png_stream_bytes(int w, int h, int color_type, int bit_depth, int interlace) {
    if (llfe_png_rowbytes(w, color_type, bit_depth) < 0 || h <= 0 || interlace < 0 || interlace > 1) return -1;
    if (!interlace) return (int64_t)h * (llfe_png_rowbytes(w, color_type, bit_depth) + 1);
    int64_t total = 0;
    for (int p = 0; p < 7; ++p) {
        const int pw = w > ADAM7[p][0] ? (w - ADAM7[p][0] + ADAM7[p][2] - 1) / ADAM7[p][2] : 0;
        const int ph = h > ADAM7[p][1] ? (h - ADAM7[p][1] + ADAM7[p][3] - 1) / ADAM7[p][3] : 0;
        if (pw && ph) total += (int64_t)ph * (llfe_png_rowbytes(pw, color_type, bit_depth) + 1);
    }
    return total;
}

extern "C" int llfe_png_reconstruct_adam7(llfe_ctx* ctx, uint8_t* d_stream, int h, int w, int color_type, int bit_depth,
                                          const uint8_t* d_palette, uint8_t* d_bgr, int32_t* d_status) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_stream != nullptr && d_bgr != nullptr && d_status != nullptr && h > 0 && h <= 65535 && w > 0);
    const int64_t total = llfe_png_stream_bytes(w, h, color_type, bit_depth, 1);
    LLFE_CHECK_ARG(total > 0 && total < 0xffffffffll && (color_type != 3 || d_palette != nullptr));
    const int bpp = png_filter_distance(color_type, bit_depth);
    LLFE_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int32_t), ctx->stream));
    PngAdam7 f;
    f.h = h, f.w = w, f.color_type = color_type, f.depth = bit_depth;
    unsigned int off = 0;
    for (int p = 0; p < 7; ++p) {
        const int pw = w > ADAM7[p][0] ? (w - ADAM7[p][0] + ADAM7[p][2] - 1) / ADAM7[p][2] : 0;
        const int ph = h > ADAM7[p][1] ? (h - ADAM7[p][1] + ADAM7[p][3] - 1) / ADAM7[p][3] : 0;
        f.off[p] = off;
        f.rowbytes[p] = 0;
        if (!pw || !ph) continue;
        const int rb = (int)llfe_png_rowbytes(pw, color_type, bit_depth);
        f.rowbytes[p] = rb;
        LLFE_TRY(launch_png_unfilter_rows(ctx, d_stream + off, 1, ph, 0, ph, rb, bpp, d_status));   // every pass is filtered on its own
        off += (unsigned int)ph * (unsigned int)(rb + 1);
    }
    LLFE_KERNEL(ctx, "k_png_adam7_to_bgr");
    k_png_adam7_to_bgr<<<dim3(ceil_div(w, 256), h), 256, 0, ctx->stream>>>(d_stream, f, d_palette, d_bgr);
    LLFE_LAUNCHED(ctx);
    return LLFE_OK;
}
