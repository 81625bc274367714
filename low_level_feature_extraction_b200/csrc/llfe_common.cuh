// Shared declarations for libllfe.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "llfe.h"

struct llfe_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    // workspace arena (grown on demand; ops carve it per call)
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // second stream + arena of llfe_analyze: the colour chain (colour pass, compaction, k-means) runs on it next to
    // the edge / shadow chain on `stream`.  While work is being enqueued for it, `stream` points at it and
    // `on_aux` selects the second arena.
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    void* ws_aux = nullptr;
    size_t ws_aux_bytes = 0;
    bool on_aux = false;
    int shadow_variant = 0;
    bool opt_serial = false;       // llfe_set_option("serial", 1): everything on one stream (for A/B timing)
    // llfe_set_option("chunk", n): images per front-kernel launch of llfe_analyze / llfe_pipeline.  The whole super-chunk
    // by default: the kernels are issue-bound, not L2-bound, so keeping a chunk's planes L2-resident (32 images) buys
    // nothing, while one launch per chain lets the colour chain reach its k-means ~2 ms earlier (6.30 -> 5.92 ms per step)
    int opt_chunk = 256;
    // pinned staging for the *_host entry points
    void* pin = nullptr;
    size_t pin_bytes = 0;
    void* dev_stage = nullptr;
    size_t dev_stage_bytes = 0;
    // INTER_AREA tables cached per (ssize, dsize)
    struct AreaTab* area_tabs = nullptr;
    // kernels whose per-device function attributes (dynamic shared memory limit, cluster size) are already set
    const void* attr_done[64] = {};
    int n_attr_done = 0;
    unsigned long long* dummy_sums = nullptr;  // sink for the shadow sums when the caller does not want them
    uint64_t launches = 0;
    // optional per-kernel CUDA-event timing (llfe_profile_begin / llfe_profile_end)
    struct ProfRec* prof = nullptr;
    int prof_cap = 0, prof_used = 0, prof_events = 0, prof_pending = -1;
    bool prof_on = false;
    // llfe_set_option: path toggles used by the parity tests (both off in production)
    bool opt_unfused = false;      // per-stage kernels instead of the fused front kernel
    int opt_inflate_threads = 4;        // llfe_set_option("inflate_threads", 1..8): decoders per deflate stream in llfe_png_decode*_host
    int opt_contour_cut_shift = 6;      // llfe_set_option("contour_cut_shift", 0..8): log2 of the rows between cut rows (tests)
    // llfe_set_option("contour_segments", v): 0 = every border followed by one thread, 1 = long borders cut into segments
    // in calls on one or two images (default), 2 = in batches as well (passes sized to 2 GB of segment tables)
    int opt_contour_segments = 1;
    bool opt_hyst_strips = false;  // multi-launch strip hysteresis instead of the cluster kernel
    bool opt_shadow_inline = false;  // adaptive threshold inside the fused front kernel instead of k_shadow
    // llfe_set_debug_buffer: validated device buffers the k-means / hysteresis kernels write phase clocks to
    unsigned long long* dbg_kmeans = nullptr;
    size_t dbg_kmeans_bytes = 0;
    unsigned long long* dbg_hyst = nullptr;
    size_t dbg_hyst_bytes = 0;
};

struct ProfRec {
    const char* name;
    cudaEvent_t a, b;
};
void llfe_prof_mark(llfe_ctx* ctx, const char* name);
void llfe_prof_stop(llfe_ctx* ctx);
// put right before a kernel launch; LLFE_LAUNCHED closes the interval
#define LLFE_KERNEL(ctx, name)                        \
    do {                                              \
        if ((ctx)->prof_on) llfe_prof_mark(ctx, name); \
    } while (0)

// true the first time `kernel` is seen on this context (function attributes are per device, contexts are per device)
static inline bool llfe_first_use(llfe_ctx* ctx, const void* kernel) {
    for (int i = 0; i < ctx->n_attr_done; ++i)
        if (ctx->attr_done[i] == kernel) return false;
    if (ctx->n_attr_done < 64) ctx->attr_done[ctx->n_attr_done++] = kernel;
    return true;
}

void llfe_set_error(const char* fmt, ...);
// h_inflate.cu: llfe_inflate_zlib with a progress mark another thread may read (bytes below it are final) and `threads`
// decoders on the one stream (speculative block starts, see there)
int llfe_inflate_zlib_progress(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len,
                               std::atomic<size_t>* progress, int threads);
int llfe_cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define LLFE_CUDA(call)                                                       \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) return llfe_cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define LLFE_CHECK_ARG(cond)                                                  \
    do {                                                                      \
        if (!(cond)) {                                                        \
            llfe_set_error("%s: invalid argument: %s", __func__, #cond);      \
            return LLFE_E_INVALID;                                            \
        }                                                                     \
    } while (0)

// First statement of every extern "C" entry point that takes a context: the calling thread may have another
// device current (a request-handler thread pool, torch's own device guard), and everything below -- workspace
// and staging allocations, kernel launches, function attributes -- must land on the context's device.  The
// caller's current device is restored on return.
struct LlfeDeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit LlfeDeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            switched = err == cudaSuccess;
        }
    }
    ~LlfeDeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};
#define LLFE_ENTER(ctx)                                                        \
    if (!(ctx)) {                                                              \
        llfe_set_error("%s: invalid argument: ctx is null", __func__);         \
        return LLFE_E_INVALID;                                                 \
    }                                                                          \
    LlfeDeviceGuard _llfe_guard((ctx)->device);                                \
    if (_llfe_guard.err != cudaSuccess) return llfe_cuda_fail(_llfe_guard.err, "cudaSetDevice", __FILE__, __LINE__)

#define LLFE_TRY(expr)                \
    do {                              \
        int _r = (expr);              \
        if (_r != LLFE_OK) return _r; \
    } while (0)

// after every kernel launch: count it and surface launch-configuration errors
#define LLFE_LAUNCHED(ctx)                      \
    do {                                        \
        (ctx)->launches++;                      \
        if ((ctx)->prof_pending >= 0) llfe_prof_stop(ctx); \
        LLFE_CUDA(cudaPeekAtLastError());       \
    } while (0)

// workspace: returns a device pointer to at least `bytes` of scratch (256-B aligned).
// May synchronise and reallocate when the arena has to grow.
int llfe_workspace(llfe_ctx* ctx, size_t bytes, void** out);

struct WsCarver {
    char* base;
    size_t off = 0;
    explicit WsCarver(void* p) : base(static_cast<char*>(p)) {}
    template <typename T>
    T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + off);
        off += (count * sizeof(T) + 255) & ~size_t(255);
        return p;
    }
    static size_t need(size_t bytes) { return (bytes + 255) & ~size_t(255); }
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t ceil_div_sz(size_t a, size_t b) { return (a + b - 1) / b; }

// ---- internal kernels' host launchers (defined across the .cu files) --------
// bit-plane layout used by the edge chain: one bit per pixel, bit (x & 31) of
// word (x >> 5); rows padded to `wpr` 32-bit words; planes of a batch back to back.
__host__ __device__ static inline int plane_wpr(int w) { return (w + 31) / 32; }

int launch_bgr2gray(llfe_ctx* ctx, const uint8_t* bgr, size_t npix, uint8_t* gray);
int launch_bgr2rgb(llfe_ctx* ctx, const uint8_t* bgr, size_t npix, uint8_t* rgb);
int launch_lut2(llfe_ctx* ctx, const uint8_t* src, size_t count, float a1, float a2, int single, uint8_t* dst);
int launch_blur5(llfe_ctx* ctx, const uint8_t* src, int n, int h, int w, int c, uint8_t* dst);
int launch_gray_blur5(llfe_ctx* ctx, const uint8_t* bgr, int n, int h, int w, uint8_t* dst);
int launch_canny_front(llfe_ctx* ctx, const uint8_t* gray, int n, int h, int w, int low, int high, uint32_t* weak,
                       uint32_t* strong);
int launch_edge_front_bgr(llfe_ctx* ctx, const uint8_t* bgr, int n, int h, int w, int low, int high, uint32_t* weak,
                          uint32_t* strong);
int launch_hysteresis(llfe_ctx* ctx, const uint32_t* weak, uint32_t* edges, int n, int h, int w, uint32_t* flags);
int launch_hysteresis_mask_cluster(llfe_ctx* ctx, const uint32_t* weak, const uint32_t* strong, int n, int h, int w,
                                   int dilate, uint8_t* mask);
int launch_plane_to_mask(llfe_ctx* ctx, const uint32_t* plane, int n, int h, int w, int dilate, uint8_t* mask);
int launch_mask_to_plane(llfe_ctx* ctx, const uint8_t* mask, const uint8_t* and_mask, int n, int h, int w,
                         uint32_t* plane);
int launch_adaptive(llfe_ctx* ctx, const uint8_t* gray, int n, int h, int w, int C, uint8_t* mask, uint64_t* sum_count);
int launch_hist256(llfe_ctx* ctx, const uint8_t* gray, int n, size_t npix_per_image, uint32_t* hist);
int launch_otsu_sweep(llfe_ctx* ctx, const uint32_t* hist, int n, size_t npix_per_image, int invert_if_light,
                      int32_t* thresh, int32_t* invert);
int launch_binarize(llfe_ctx* ctx, const uint8_t* gray, int n, size_t npix_per_image, const int32_t* thresh,
                    const int32_t* invert, uint8_t* mask);
size_t hysteresis_flag_words(int n, int h);
int launch_dilate3_u8(llfe_ctx* ctx, const uint8_t* src, int n, int h, int w, uint8_t* dst);
void llfe_free_area_tabs(llfe_ctx* ctx);
int launch_unique_colors(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, const int8_t* d_noise, uint64_t seed,
                         int first_image, uint32_t* d_keys, uint32_t* d_hist, int32_t* d_count, int max_unique);

size_t bitmap_words_per_image();
size_t bitmap_blocks_per_image();
int launch_bitmap_compact(llfe_ctx* ctx, const uint32_t* bitmap, uint32_t* bsum, int m, uint32_t* d_keys, uint32_t* rank,
                          int32_t* d_count, int max_unique);
bool fused_supported(int h, int w);
int launch_fused(llfe_ctx* ctx, const uint8_t* bgr, int n, int h, int w, int low, int high, uint32_t* weak,
                 uint32_t* strong, uint8_t* mask, uint64_t* sum_count, uint8_t* blur_out);
int launch_color_bitmap(llfe_ctx* ctx, const uint8_t* d_bgr, int m, int h, int w, const int8_t* d_noise, uint64_t seed,
                        int first_image, uint32_t* bitmap);
bool shadow_split_supported(int h, int w);
// JPEG (k_jpeg.cu): entropy decoding on the host into `pin`, the rest on the device
size_t llfe_jpeg_stage_bytes(int h, int w);
int llfe_jpeg_decode_impl(llfe_ctx* ctx, const uint8_t* buf, size_t len, int h, int w, uint8_t* h_bgr, uint8_t* pin, uint8_t* dev,
                          size_t cap);
// PNG reconstruction (k_png.cu): rows [row0, row1) in place, then the conversion to BGR
int png_filter_distance(int color_type, int bit_depth);
int launch_png_unfilter_rows(llfe_ctx* ctx, uint8_t* d_stream, int n, int h, int row0, int row1, int rowbytes, int bpp,
                             int32_t* d_status);
int launch_png_to_bgr(llfe_ctx* ctx, const uint8_t* d_stream, int n, int h, int w, int rowbytes, int color_type, int bit_depth,
                      const uint8_t* d_palette, uint8_t* d_bgr);
int launch_shadow(llfe_ctx* ctx, const uint8_t* blurred, int n, int h, int w, uint8_t* mask, uint64_t* sum_count);
