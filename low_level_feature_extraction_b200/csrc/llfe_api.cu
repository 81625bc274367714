// extern "C" surface of libllfe.so: context, memory helpers and the per-op
// entry points declared in include/llfe.h.  Kernels live in the k_*.cu files.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <atomic>
#include <thread>
#include <vector>

#include "llfe_common.cuh"

#define LLFE_VERSION_NUM 100

static thread_local char g_err[512] = "";

void llfe_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int llfe_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    llfe_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return e == cudaErrorMemoryAllocation ? LLFE_E_NOMEM : LLFE_E_CUDA;
}

int llfe_workspace(llfe_ctx* ctx, size_t bytes, void** out) {
    void*& ws = ctx->on_aux ? ctx->ws_aux : ctx->ws;
    size_t& have = ctx->on_aux ? ctx->ws_aux_bytes : ctx->ws_bytes;
    if (bytes > have) {
        // grow: everything queued so far on this arena's stream may still be using the old arena
        LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ws) LLFE_CUDA(cudaFree(ws));
        ws = nullptr;
        have = 0;
        size_t want = bytes + bytes / 4 + (1 << 20);
        LLFE_CUDA(cudaMalloc(&ws, want));
        have = want;
    }
    *out = ws;
    return LLFE_OK;
}

// Second stream of the context (created on first use) and the fork / join around work enqueued on it.
static int aux_begin(llfe_ctx* ctx, cudaStream_t* main_out) {
    if (!ctx->aux_stream) {
        LLFE_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        LLFE_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        LLFE_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
    *main_out = ctx->stream;
    LLFE_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    LLFE_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    return LLFE_OK;
}
static inline void aux_enter(llfe_ctx* ctx) {
    ctx->stream = ctx->aux_stream;
    ctx->on_aux = true;
}
static inline void aux_leave(llfe_ctx* ctx, cudaStream_t main_stream) {
    ctx->stream = main_stream;
    ctx->on_aux = false;
}
static int aux_join(llfe_ctx* ctx) {
    LLFE_CUDA(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
    LLFE_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    return LLFE_OK;
}

// ---- per-kernel CUDA-event timing ---------------------------------------------------
void llfe_prof_mark(llfe_ctx* ctx, const char* name) {
    if (!ctx->prof_on || ctx->prof_used >= ctx->prof_cap) return;
    int i = ctx->prof_used;
    if (i >= ctx->prof_events) {
        if (cudaEventCreate(&ctx->prof[i].a) != cudaSuccess || cudaEventCreate(&ctx->prof[i].b) != cudaSuccess) return;
        ctx->prof_events = i + 1;
    }
    ctx->prof[i].name = name;
    cudaEventRecord(ctx->prof[i].a, ctx->stream);
    ctx->prof_pending = i;
    ctx->prof_used = i + 1;
}

void llfe_prof_stop(llfe_ctx* ctx) {
    if (ctx->prof_pending >= 0) cudaEventRecord(ctx->prof[ctx->prof_pending].b, ctx->stream);
    ctx->prof_pending = -1;
}

static int ensure_stage(llfe_ctx* ctx, size_t pin_bytes, size_t dev_bytes) {
    if (pin_bytes > ctx->pin_bytes) {
        LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->pin) LLFE_CUDA(cudaFreeHost(ctx->pin));
        ctx->pin = nullptr;
        ctx->pin_bytes = 0;
        LLFE_CUDA(cudaMallocHost(&ctx->pin, pin_bytes));
        ctx->pin_bytes = pin_bytes;
    }
    if (dev_bytes > ctx->dev_stage_bytes) {
        LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->dev_stage) LLFE_CUDA(cudaFree(ctx->dev_stage));
        ctx->dev_stage = nullptr;
        ctx->dev_stage_bytes = 0;
        LLFE_CUDA(cudaMalloc(&ctx->dev_stage, dev_bytes));
        ctx->dev_stage_bytes = dev_bytes;
    }
    return LLFE_OK;
}

extern "C" {

int llfe_version(void) { return LLFE_VERSION_NUM; }
const char* llfe_last_error(void) { return g_err; }

int llfe_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int llfe_create(int device, llfe_ctx** out) {
    LLFE_CHECK_ARG(out != nullptr);
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        llfe_set_error("llfe_create: no usable CUDA device (%s); this library has no CPU fallback",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return LLFE_E_NODEVICE;
    }
    LLFE_CHECK_ARG(device >= 0 && device < n);
    LLFE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LLFE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        llfe_set_error("llfe_create: device %d is sm_%d%d; libllfe.so carries sm_100a code only", device, prop.major,
                       prop.minor);
        return LLFE_E_UNSUPPORTED;
    }
    llfe_ctx* c = new (std::nothrow) llfe_ctx();
    if (!c) return LLFE_E_NOMEM;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    cudaError_t se = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) {
        delete c;
        return llfe_cuda_fail(se, "cudaStreamCreateWithFlags", __FILE__, __LINE__);
    }
    c->stream = c->own_stream;
    *out = c;
    return LLFE_OK;
}

int llfe_destroy(llfe_ctx* ctx) {
    if (!ctx) return LLFE_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->ws_aux) cudaFree(ctx->ws_aux);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->pin) cudaFreeHost(ctx->pin);
    if (ctx->dev_stage) cudaFree(ctx->dev_stage);
    if (ctx->dummy_sums) cudaFree(ctx->dummy_sums);
    llfe_free_area_tabs(ctx);
    if (ctx->prof) {
        for (int i = 0; i < ctx->prof_events; ++i) {
            cudaEventDestroy(ctx->prof[i].a);
            cudaEventDestroy(ctx->prof[i].b);
        }
        delete[] ctx->prof;
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return LLFE_OK;
}

int llfe_set_stream(llfe_ctx* ctx, void* cuda_stream) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr);
    ctx->stream = (cudaStream_t)cuda_stream;
    return LLFE_OK;
}

int llfe_use_own_stream(llfe_ctx* ctx) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr);
    ctx->stream = ctx->own_stream;
    return LLFE_OK;
}

int llfe_set_option(llfe_ctx* ctx, const char* name, int64_t value) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(name != nullptr);
    if (!strcmp(name, "unfused")) ctx->opt_unfused = value != 0;
    else if (!strcmp(name, "hyst_strips")) ctx->opt_hyst_strips = value != 0;
    else if (!strcmp(name, "shadow_inline")) ctx->opt_shadow_inline = value != 0;
    else if (!strcmp(name, "serial")) ctx->opt_serial = value != 0;
    else if (!strcmp(name, "inflate_threads")) ctx->opt_inflate_threads = (value >= 1 && value <= 8) ? (int)value : 4;
    else if (!strcmp(name, "contour_segments")) ctx->opt_contour_segments = (value >= 0 && value <= 2) ? (int)value : 1;
    else if (!strcmp(name, "contour_cut_shift")) ctx->opt_contour_cut_shift = (value >= 0 && value <= 8) ? (int)value : 6;
    else if (!strcmp(name, "shadow_variant")) ctx->shadow_variant = (int)value;
    else if (!strcmp(name, "chunk")) ctx->opt_chunk = (value >= 1 && value <= 256) ? (int)value : 256;
    else {
        llfe_set_error("llfe_set_option: unknown option '%s'", name);
        return LLFE_E_INVALID;
    }
    return LLFE_OK;
}

int llfe_set_debug_buffer(llfe_ctx* ctx, const char* name, void* d_buf, size_t bytes) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(name != nullptr);
    if (d_buf) {   // must be device memory of this context's device, at least `bytes` long as far as CUDA can tell
        cudaPointerAttributes at;
        cudaError_t e = cudaPointerGetAttributes(&at, d_buf);
        if (e != cudaSuccess || at.type != cudaMemoryTypeDevice || at.device != ctx->device) {
            cudaGetLastError();
            llfe_set_error("llfe_set_debug_buffer: %p is not device memory of device %d", d_buf, ctx->device);
            return LLFE_E_INVALID;
        }
    } else {
        bytes = 0;
    }
    if (!strcmp(name, "kmeans")) {
        ctx->dbg_kmeans = (unsigned long long*)d_buf;
        ctx->dbg_kmeans_bytes = bytes;
    } else if (!strcmp(name, "hysteresis")) {
        ctx->dbg_hyst = (unsigned long long*)d_buf;
        ctx->dbg_hyst_bytes = bytes;
    } else {
        llfe_set_error("llfe_set_debug_buffer: unknown buffer '%s'", name);
        return LLFE_E_INVALID;
    }
    return LLFE_OK;
}

int llfe_sync(llfe_ctx* ctx) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr);
    LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
    return LLFE_OK;
}

int llfe_profile_begin(llfe_ctx* ctx) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr);
    if (!ctx->prof) {
        ctx->prof_cap = 1 << 16;
        ctx->prof = new (std::nothrow) ProfRec[ctx->prof_cap];
        if (!ctx->prof) return LLFE_E_NOMEM;
    }
    ctx->prof_used = 0;
    ctx->prof_pending = -1;
    ctx->prof_on = true;
    return LLFE_OK;
}

int llfe_profile_end(llfe_ctx* ctx, char* json, size_t cap) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && json != nullptr && cap >= 64);
    ctx->prof_on = false;
    LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
    // aggregate by name (names are string literals: compare by content, few distinct)
    struct Agg {
        const char* name;
        double ms;
        long n;
    } agg[64];
    int na = 0;
    for (int i = 0; i < ctx->prof_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->prof[i].a, ctx->prof[i].b) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        int k = 0;
        for (; k < na; ++k)
            if (strcmp(agg[k].name, ctx->prof[i].name) == 0) break;
        if (k == na) {
            if (na == 64) continue;
            agg[na++] = {ctx->prof[i].name, 0.0, 0};
        }
        agg[k].ms += ms;
        agg[k].n += 1;
    }
    size_t off = 0;
    off += snprintf(json + off, cap - off, "{");
    for (int k = 0; k < na && off + 96 < cap; ++k)
        off += snprintf(json + off, cap - off, "%s\"%s\": {\"ms\": %.6f, \"launches\": %ld}", k ? ", " : "", agg[k].name,
                        agg[k].ms, agg[k].n);
    snprintf(json + off, cap - off, "}");
    return LLFE_OK;
}

uint64_t llfe_launch_count(llfe_ctx* ctx) { return ctx ? ctx->launches : 0; }
int llfe_sm_count(llfe_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

int llfe_malloc(llfe_ctx* ctx, size_t bytes, void** d_out) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_out != nullptr);
    LLFE_CUDA(cudaMalloc(d_out, bytes ? bytes : 1));
    return LLFE_OK;
}
int llfe_free(llfe_ctx* ctx, void* d_ptr) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr);
    if (d_ptr) {
        LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
        LLFE_CUDA(cudaFree(d_ptr));
    }
    return LLFE_OK;
}
int llfe_malloc_host(llfe_ctx* ctx, size_t bytes, void** h_out) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && h_out != nullptr);
    LLFE_CUDA(cudaMallocHost(h_out, bytes ? bytes : 1));
    return LLFE_OK;
}
int llfe_free_host(llfe_ctx* ctx, void* h_ptr) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr);
    if (h_ptr) LLFE_CUDA(cudaFreeHost(h_ptr));
    return LLFE_OK;
}
int llfe_memcpy_h2d(llfe_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr);
    if (bytes) LLFE_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return LLFE_OK;
}
int llfe_memcpy_d2h(llfe_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr);
    if (bytes) LLFE_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return LLFE_OK;
}
int llfe_memset(llfe_ctx* ctx, void* d_dst, int value, size_t bytes) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr);
    if (bytes) LLFE_CUDA(cudaMemsetAsync(d_dst, value, bytes, ctx->stream));
    return LLFE_OK;
}

#define LLFE_IMG_ARGS(ptr) LLFE_CHECK_ARG(ctx != nullptr && (ptr) != nullptr && n >= 0 && h >= 0 && w >= 0 && n <= 65535)

// ---- pointwise ----------------------------------------------------------------
int llfe_bgr2gray(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, uint8_t* d_gray) {
    LLFE_ENTER(ctx);
    LLFE_IMG_ARGS(d_bgr);
    LLFE_CHECK_ARG(d_gray != nullptr);
    return launch_bgr2gray(ctx, d_bgr, (size_t)n * h * w, d_gray);
}

int llfe_bgr2rgb(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, uint8_t* d_rgb) {
    LLFE_ENTER(ctx);
    LLFE_IMG_ARGS(d_bgr);
    LLFE_CHECK_ARG(d_rgb != nullptr);
    return launch_bgr2rgb(ctx, d_bgr, (size_t)n * h * w, d_rgb);
}

int llfe_convert_scale_abs(llfe_ctx* ctx, const uint8_t* d_src, size_t count, float a1, float a2, int single,
                           uint8_t* d_dst) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && d_src != nullptr && d_dst != nullptr);
    return launch_lut2(ctx, d_src, count, a1, a2, single, d_dst);
}

// ---- blur -------------------------------------------------------------------
int llfe_gaussian_blur5(llfe_ctx* ctx, const uint8_t* d_src, int n, int h, int w, int c, uint8_t* d_dst) {
    LLFE_ENTER(ctx);
    LLFE_IMG_ARGS(d_src);
    LLFE_CHECK_ARG(d_dst != nullptr && (c == 1 || c == 3) && d_src != d_dst);
    return launch_blur5(ctx, d_src, n, h, w, c, d_dst);
}

int llfe_gray_blur5(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, uint8_t* d_blurred) {
    LLFE_ENTER(ctx);
    LLFE_IMG_ARGS(d_bgr);
    LLFE_CHECK_ARG(d_blurred != nullptr);
    return launch_gray_blur5(ctx, d_bgr, n, h, w, d_blurred);
}

// ---- edges ------------------------------------------------------------------
// weak + strong planes -> (dilated) u8 mask: one cluster launch per batch when the strips of an image
// fit in shared memory, else the multi-launch strip kernels + the expand kernel.
static int hysteresis_to_mask(llfe_ctx* ctx, const uint32_t* weak, uint32_t* edges, int n, int h, int w,
                              uint32_t* flags, int dilate, uint8_t* d_out) {
    if (!ctx->opt_hyst_strips) {
        const int rc = launch_hysteresis_mask_cluster(ctx, weak, edges, n, h, w, dilate, d_out);
        if (rc != LLFE_E_UNSUPPORTED) return rc;
    }
    LLFE_TRY(launch_hysteresis(ctx, weak, edges, n, h, w, flags));
    return launch_plane_to_mask(ctx, edges, n, h, w, dilate, d_out);
}

static int canny_from_gray(llfe_ctx* ctx, const uint8_t* d_gray, int n, int h, int w, int low, int high, int dilate,
                           uint8_t* d_out, char* ws_base) {
    const size_t plane = (size_t)n * h * plane_wpr(w);
    WsCarver ws(ws_base);
    uint32_t* weak = ws.take<uint32_t>(plane);
    uint32_t* edges = ws.take<uint32_t>(plane);
    uint32_t* flags = ws.take<uint32_t>(hysteresis_flag_words(n, h));
    LLFE_TRY(launch_canny_front(ctx, d_gray, n, h, w, low, high, weak, edges));
    return hysteresis_to_mask(ctx, weak, edges, n, h, w, flags, dilate, d_out);
}

static size_t canny_ws_bytes(int n, int h, int w) {
    const size_t plane = (size_t)n * h * plane_wpr(w) * sizeof(uint32_t);
    return 2 * WsCarver::need(plane) + WsCarver::need(hysteresis_flag_words(n, h) * sizeof(uint32_t));
}

int llfe_canny(llfe_ctx* ctx, const uint8_t* d_gray, int n, int h, int w, int low, int high, uint8_t* d_edges) {
    LLFE_ENTER(ctx);
    LLFE_IMG_ARGS(d_gray);
    LLFE_CHECK_ARG(d_edges != nullptr);
    if ((size_t)n * h * w == 0) return LLFE_OK;
    void* ws;
    LLFE_TRY(llfe_workspace(ctx, canny_ws_bytes(n, h, w), &ws));
    return canny_from_gray(ctx, d_gray, n, h, w, low, high, 0, d_edges, (char*)ws);
}

int llfe_hysteresis(llfe_ctx* ctx, const uint8_t* d_weak, const uint8_t* d_strong, int n, int h, int w, int dilate,
                    uint8_t* d_edges) {
    LLFE_ENTER(ctx);
    LLFE_IMG_ARGS(d_weak);
    LLFE_CHECK_ARG(d_strong != nullptr && d_edges != nullptr);
    if ((size_t)n * h * w == 0) return LLFE_OK;
    void* wsp;
    LLFE_TRY(llfe_workspace(ctx, canny_ws_bytes(n, h, w), &wsp));
    const size_t plane = (size_t)n * h * plane_wpr(w);
    WsCarver ws(wsp);
    uint32_t* weak = ws.take<uint32_t>(plane);
    uint32_t* edges = ws.take<uint32_t>(plane);
    uint32_t* flags = ws.take<uint32_t>(hysteresis_flag_words(n, h));
    LLFE_TRY(launch_mask_to_plane(ctx, d_weak, nullptr, n, h, w, weak));
    LLFE_TRY(launch_mask_to_plane(ctx, d_strong, d_weak, n, h, w, edges));   // seeds outside the weak set cannot grow
    return hysteresis_to_mask(ctx, weak, edges, n, h, w, flags, dilate, d_edges);
}

int llfe_dilate3(llfe_ctx* ctx, const uint8_t* d_src, int n, int h, int w, uint8_t* d_dst) {
    LLFE_ENTER(ctx);
    LLFE_IMG_ARGS(d_src);
    LLFE_CHECK_ARG(d_dst != nullptr && d_src != d_dst);
    return launch_dilate3_u8(ctx, d_src, n, h, w, d_dst);
}

int llfe_shape_mask(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, int low, int high, uint8_t* d_mask) {
    LLFE_ENTER(ctx);
    LLFE_IMG_ARGS(d_bgr);
    LLFE_CHECK_ARG(d_mask != nullptr);
    if ((size_t)n * h * w == 0) return LLFE_OK;
    if (fused_supported(h, w) && !ctx->opt_unfused)
        return llfe_pipeline(ctx, d_bgr, n, h, w, low, high, d_mask, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0);
    const size_t img = WsCarver::need((size_t)n * h * w);
    void* ws;
    LLFE_TRY(llfe_workspace(ctx, img + canny_ws_bytes(n, h, w), &ws));
    uint8_t* blurred = (uint8_t*)ws;
    LLFE_TRY(launch_gray_blur5(ctx, d_bgr, n, h, w, blurred));
    return canny_from_gray(ctx, blurred, n, h, w, low, high, 1, d_mask, (char*)ws + img);
}

// ---- thresholds ---------------------------------------------------------------
int llfe_adaptive_threshold(llfe_ctx* ctx, const uint8_t* d_gray, int n, int h, int w, int C, uint8_t* d_mask,
                            uint64_t* d_sum_count) {
    LLFE_ENTER(ctx);
    LLFE_IMG_ARGS(d_gray);
    LLFE_CHECK_ARG(d_mask != nullptr);
    return launch_adaptive(ctx, d_gray, n, h, w, C, d_mask, d_sum_count);
}

int llfe_shadow_mask(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, uint8_t* d_mask, uint8_t* d_blurred,
                     uint64_t* d_sum_count) {
    LLFE_ENTER(ctx);
    LLFE_IMG_ARGS(d_bgr);
    LLFE_CHECK_ARG(d_mask != nullptr);
    if ((size_t)n * h * w == 0) return LLFE_OK;
    if (!d_blurred && fused_supported(h, w) && !ctx->opt_unfused)
        return llfe_pipeline(ctx, d_bgr, n, h, w, 50, 150, nullptr, d_mask, d_sum_count, nullptr, 0, nullptr, nullptr, 0);
    uint8_t* blurred = d_blurred;
    if (!blurred) {
        void* ws;
        LLFE_TRY(llfe_workspace(ctx, (size_t)n * h * w, &ws));
        blurred = (uint8_t*)ws;
    }
    LLFE_TRY(launch_gray_blur5(ctx, d_bgr, n, h, w, blurred));
    return launch_adaptive(ctx, blurred, n, h, w, 2, d_mask, d_sum_count);
}

int llfe_font_mask(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, uint8_t* d_mask) {
    LLFE_ENTER(ctx);
    LLFE_IMG_ARGS(d_bgr);
    LLFE_CHECK_ARG(d_mask != nullptr);
    if ((size_t)n * h * w == 0) return LLFE_OK;
    void* ws;
    LLFE_TRY(llfe_workspace(ctx, (size_t)n * h * w, &ws));
    LLFE_TRY(launch_bgr2gray(ctx, d_bgr, (size_t)n * h * w, (uint8_t*)ws));
    return launch_adaptive(ctx, (uint8_t*)ws, n, h, w, 2, d_mask, nullptr);
}

static int otsu_impl(llfe_ctx* ctx, const uint8_t* d_gray, int n, size_t npix, int invert_if_light, uint8_t* d_mask,
                     int32_t* d_thresh, char* ws_base) {
    WsCarver ws(ws_base);
    uint32_t* hist = ws.take<uint32_t>((size_t)n * 256);
    int32_t* thr = ws.take<int32_t>(n);
    int32_t* inv = ws.take<int32_t>(n);
    LLFE_TRY(launch_hist256(ctx, d_gray, n, npix, hist));
    LLFE_TRY(launch_otsu_sweep(ctx, hist, n, npix, invert_if_light, d_thresh ? d_thresh : thr, inv));
    return launch_binarize(ctx, d_gray, n, npix, d_thresh ? d_thresh : thr, inv, d_mask);
}

static size_t otsu_ws_bytes(int n) {
    return WsCarver::need((size_t)n * 256 * 4) + 2 * WsCarver::need((size_t)n * 4);
}

int llfe_otsu(llfe_ctx* ctx, const uint8_t* d_gray, int n, int h, int w, int invert_if_light, uint8_t* d_mask,
              int32_t* d_thresh) {
    LLFE_ENTER(ctx);
    LLFE_IMG_ARGS(d_gray);
    LLFE_CHECK_ARG(d_mask != nullptr);
    if ((size_t)n * h * w == 0) return LLFE_OK;
    void* ws;
    LLFE_TRY(llfe_workspace(ctx, otsu_ws_bytes(n), &ws));
    return otsu_impl(ctx, d_gray, n, (size_t)h * w, invert_if_light, d_mask, d_thresh, (char*)ws);
}

int llfe_text_mask(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, uint8_t* d_mask, int32_t* d_thresh) {
    LLFE_ENTER(ctx);
    LLFE_IMG_ARGS(d_bgr);
    LLFE_CHECK_ARG(d_mask != nullptr);
    if ((size_t)n * h * w == 0) return LLFE_OK;
    const size_t img = WsCarver::need((size_t)n * h * w);
    void* ws;
    LLFE_TRY(llfe_workspace(ctx, img + otsu_ws_bytes(n), &ws));
    uint8_t* gray = (uint8_t*)ws;
    LLFE_TRY(launch_bgr2gray(ctx, d_bgr, (size_t)n * h * w, gray));
    return otsu_impl(ctx, gray, n, (size_t)h * w, 1, d_mask, d_thresh, (char*)ws + img);
}

// ---- fused service pipeline ------------------------------------------------------
// k-means arguments of llfe_analyze (null = llfe_pipeline: stop after the unique-colour lists)
struct KmeansCall {
    int k, attempts, max_iter;
    double eps;
    const uint64_t* rng_state;
    float* centers;
    int32_t* labels;
    int32_t* k_used;
    int32_t* sizes;
    int32_t* status;
};

static int analyze_impl(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, int low, int high, uint8_t* d_shape_mask,
                        uint8_t* d_shadow_mask, uint64_t* d_shadow_sum_count, const int8_t* d_noise, uint64_t seed,
                        uint32_t* d_keys, int32_t* d_count, int max_unique, const KmeansCall* km) {
    LLFE_IMG_ARGS(d_bgr);
    LLFE_CHECK_ARG(h > 0 && w > 0);
    LLFE_CHECK_ARG(d_keys == nullptr || (d_count != nullptr && max_unique > 0));
    if (n == 0) return LLFE_OK;
    if (fused_supported(h, w) && !ctx->opt_unfused) {
        // Per chunk (by default the whole super-chunk of up to 256 images): the front kernel (edge bit planes + the blurred
        // gray plane), the shadow kernel on that plane, the colour pass and the ordered compaction.  The hysteresis is
        // latency-bound (a few busy warps per image), so it runs once per super-chunk to have as many images in flight as
        // the SMs can hold.
        const int chunk = n < ctx->opt_chunk ? n : ctx->opt_chunk;
        const int super = n < 256 ? n : 256;
        const size_t plane_words = (size_t)super * h * plane_wpr(w), plane_img = (size_t)h * plane_wpr(w);
        const size_t bmw = bitmap_words_per_image(), bmb = bitmap_blocks_per_image();
        const size_t p = (size_t)h * w;
        // The adaptive threshold runs in its own kernel (k_shadow) on the blurred plane the front kernel writes for
        // the chunk (L2-sized); "shadow_inline" keeps it inside the front kernel (the round-1 layout, for comparison).
        const bool split = d_shadow_mask && !ctx->opt_shadow_inline;
        const bool front = d_shape_mask != nullptr;   // anything for the front kernel besides the blurred plane?
        const bool masks = d_shape_mask || d_shadow_mask, colours = d_keys != nullptr;
        // Two independent chains: masks (front kernel, shadow kernel, hysteresis) and colours (colour pass, compaction,
        // k-means).  With both present they run on two streams: the latency-bound kernels of one chain (hysteresis,
        // compaction, the k-means tail) fill issue slots the other leaves idle.  Results do not depend on the schedule.
        const bool two = masks && colours && !ctx->opt_serial;
        const size_t need_colour = WsCarver::need(bmw * 4 * chunk) + WsCarver::need(bmb * 4 * chunk);
        size_t need = 2 * WsCarver::need(plane_words * 4) + WsCarver::need(hysteresis_flag_words(super, h) * 4);
        if (colours && !two) need += need_colour;
        if (split) need += WsCarver::need(p * chunk);
        void* ws;
        LLFE_TRY(llfe_workspace(ctx, need, &ws));
        WsCarver carve(ws);
        uint32_t* weak = carve.take<uint32_t>(plane_words);
        uint32_t* edges = carve.take<uint32_t>(plane_words);
        uint32_t* flags = carve.take<uint32_t>(hysteresis_flag_words(super, h));
        uint8_t* blurred = split ? carve.take<uint8_t>(p * chunk) : nullptr;
        uint32_t* bitmap = (colours && !two) ? carve.take<uint32_t>(bmw * chunk) : nullptr;
        uint32_t* bsum = (colours && !two) ? carve.take<uint32_t>(bmb * chunk) : nullptr;
        cudaStream_t main_stream = ctx->stream;
        struct AuxScope {   // whatever happens below, the context goes back to its own stream
            llfe_ctx* c;
            cudaStream_t s;
            ~AuxScope() {
                c->stream = s;
                c->on_aux = false;
            }
        } scope{ctx, main_stream};
        if (two) LLFE_TRY(aux_begin(ctx, &main_stream));
        for (int s0 = 0; s0 < n; s0 += super) {
            const int sm = (n - s0) < super ? (n - s0) : super;
            if (two) {   // the colour chain's scratch lives in the second arena (k-means may have grown it: carve again)
                aux_enter(ctx);
                void* wsc;
                LLFE_TRY(llfe_workspace(ctx, need_colour, &wsc));
                WsCarver cc(wsc);
                bitmap = cc.take<uint32_t>(bmw * chunk);
                bsum = cc.take<uint32_t>(bmb * chunk);
                aux_leave(ctx, main_stream);
            }
            for (int i0 = s0; i0 < s0 + sm; i0 += chunk) {
                const int m = (s0 + sm - i0) < chunk ? (s0 + sm - i0) : chunk;
                if (masks) {
                    if (split && !front) {
                        LLFE_TRY(launch_gray_blur5(ctx, d_bgr + i0 * p * 3, m, h, w, blurred));
                    } else {
                        LLFE_TRY(launch_fused(ctx, d_bgr + i0 * p * 3, m, h, w, low, high,
                                              d_shape_mask ? weak + (i0 - s0) * plane_img : nullptr,
                                              d_shape_mask ? edges + (i0 - s0) * plane_img : nullptr,
                                              (d_shadow_mask && !split) ? d_shadow_mask + i0 * p : nullptr,
                                              (d_shadow_sum_count && !split) ? d_shadow_sum_count + 2 * i0 : nullptr, blurred));
                    }
                    if (split)
                        LLFE_TRY(launch_shadow(ctx, blurred, m, h, w, d_shadow_mask + i0 * p,
                                               d_shadow_sum_count ? d_shadow_sum_count + 2 * i0 : nullptr));
                }
                if (colours) {
                    // a pointwise pass of its own over the chunk (bitmaps of 32 images stay in L2), then the ordered compaction
                    if (two) aux_enter(ctx);
                    LLFE_CUDA(cudaMemsetAsync(bitmap, 0, bmw * 4 * m, ctx->stream));
                    LLFE_TRY(launch_color_bitmap(ctx, d_bgr + i0 * p * 3, m, h, w, d_noise ? d_noise + i0 * p * 3 : nullptr, seed,
                                                 i0, bitmap));
                    LLFE_TRY(launch_bitmap_compact(ctx, bitmap, bsum, m, d_keys + (size_t)i0 * max_unique, nullptr,
                                                   d_count + i0, max_unique));
                    if (two) aux_leave(ctx, main_stream);
                }
            }
            if (d_shape_mask) LLFE_TRY(hysteresis_to_mask(ctx, weak, edges, sm, h, w, flags, 1, d_shape_mask + s0 * p));
            if (km) {
                if (two) aux_enter(ctx);
                LLFE_TRY(llfe_kmeans_unique(ctx, d_keys + (size_t)s0 * max_unique, d_count + s0, sm, max_unique, km->k, km->attempts,
                                            km->max_iter, km->eps, km->rng_state + s0, km->centers + (size_t)s0 * km->k * 3,
                                            km->labels ? km->labels + (size_t)s0 * max_unique : nullptr, nullptr,
                                            km->k_used ? km->k_used + s0 : nullptr, km->sizes ? km->sizes + (size_t)s0 * km->k : nullptr,
                                            km->status ? km->status + s0 : nullptr));
                if (two) aux_leave(ctx, main_stream);
            }
        }
        if (two) LLFE_TRY(aux_join(ctx));
        return LLFE_OK;
    }
    if (d_shape_mask || d_shadow_mask) {
        // gray + blur once, shared by the edge chain and the adaptive threshold
        const size_t img = WsCarver::need((size_t)n * h * w);
        void* ws;
        LLFE_TRY(llfe_workspace(ctx, img + canny_ws_bytes(n, h, w), &ws));
        uint8_t* blurred = (uint8_t*)ws;
        LLFE_TRY(launch_gray_blur5(ctx, d_bgr, n, h, w, blurred));
        if (d_shape_mask) LLFE_TRY(canny_from_gray(ctx, blurred, n, h, w, low, high, 1, d_shape_mask, (char*)ws + img));
        if (d_shadow_mask) LLFE_TRY(launch_adaptive(ctx, blurred, n, h, w, 2, d_shadow_mask, d_shadow_sum_count));
    }
    if (d_keys) LLFE_TRY(launch_unique_colors(ctx, d_bgr, n, h, w, d_noise, seed, 0, d_keys, nullptr, d_count, max_unique));
    if (km)
        LLFE_TRY(llfe_kmeans_unique(ctx, d_keys, d_count, n, max_unique, km->k, km->attempts, km->max_iter, km->eps, km->rng_state,
                                    km->centers, km->labels, nullptr, km->k_used, km->sizes, km->status));
    return LLFE_OK;
}

int llfe_pipeline(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, int low, int high, uint8_t* d_shape_mask,
                  uint8_t* d_shadow_mask, uint64_t* d_shadow_sum_count, const int8_t* d_noise, uint64_t seed,
                  uint32_t* d_keys, int32_t* d_count, int max_unique) {
    LLFE_ENTER(ctx);
    return analyze_impl(ctx, d_bgr, n, h, w, low, high, d_shape_mask, d_shadow_mask, d_shadow_sum_count, d_noise, seed, d_keys,
                        d_count, max_unique, nullptr);
}

int llfe_analyze(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, int low, int high, uint8_t* d_shape_mask,
                 uint8_t* d_shadow_mask, uint64_t* d_shadow_sum_count, const int8_t* d_noise, uint64_t seed, uint32_t* d_keys,
                 int32_t* d_count, int max_unique, int k, int attempts, int max_iter, double eps, const uint64_t* d_rng_state,
                 float* d_centers, int32_t* d_labels, int32_t* d_k_used, int32_t* d_cluster_sizes, int32_t* d_status) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_keys != nullptr && d_count != nullptr && d_rng_state != nullptr && d_centers != nullptr);
    KmeansCall km{k, attempts, max_iter, eps, d_rng_state, d_centers, d_labels, d_k_used, d_cluster_sizes, d_status};
    return analyze_impl(ctx, d_bgr, n, h, w, low, high, d_shape_mask, d_shadow_mask, d_shadow_sum_count, d_noise, seed, d_keys,
                        d_count, max_unique, &km);
}

// ---- bit-packed masks for the host path ---------------------------------------------------
// A u8 mask in {0, 255} carries one bit per pixel: over PCIe it travels as a bit plane (P/8 bytes instead of P) and is
// expanded to the reference's u8 array by host threads.  Plane layout as everywhere in the library: bit (x & 31) of word
// (x >> 5), rows padded to plane_wpr(w) words, images back to back.
int llfe_pack_mask_bits(llfe_ctx* ctx, const uint8_t* d_mask, int n, int h, int w, uint32_t* d_bits) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(d_mask != nullptr && d_bits != nullptr && n >= 0 && h >= 0 && w >= 0 && n <= 65535);
    return launch_mask_to_plane(ctx, d_mask, nullptr, n, h, w, d_bits);
}

int llfe_mask_bits_words_per_row(int w) { return plane_wpr(w); }

int llfe_expand_mask_bits_host(const uint32_t* h_bits, int n, int h, int w, uint8_t* h_mask, int threads) {
    if (!h_bits || !h_mask || n < 0 || h < 0 || w < 0) {
        llfe_set_error("llfe_expand_mask_bits_host: invalid argument");
        return LLFE_E_INVALID;
    }
    static uint64_t lut[256];
    static bool lut_ready = false;
    if (!lut_ready) {   // byte of 8 mask bits -> 8 mask bytes (benign race: every thread writes the same values)
        for (int b = 0; b < 256; ++b) {
            uint64_t v = 0;
            for (int j = 0; j < 8; ++j)
                if (b & (1 << j)) v |= 0xffull << (8 * j);
            lut[b] = v;
        }
        lut_ready = true;
    }
    const int wpr = plane_wpr(w);
    const size_t rows = (size_t)n * h;
    auto work = [=](size_t r0, size_t r1) {
        for (size_t r = r0; r < r1; ++r) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(h_bits + r * wpr);
            uint8_t* dst = h_mask + r * (size_t)w;
            const int full = w >> 3;
            for (int i = 0; i < full; ++i) {
                const uint64_t v = lut[src[i]];
                memcpy(dst + 8 * (size_t)i, &v, 8);
            }
            for (int x = full << 3; x < w; ++x) dst[x] = (src[x >> 3] >> (x & 7)) & 1 ? 255 : 0;
        }
    };
    int t = threads < 1 ? 1 : threads;
    if ((size_t)t > rows) t = rows ? (int)rows : 1;
    if (t == 1) {
        work(0, rows);
        return LLFE_OK;
    }
    std::vector<std::thread> pool;
    for (int i = 0; i < t; ++i) pool.emplace_back(work, rows * i / t, rows * (i + 1) / t);
    for (auto& th : pool) th.join();
    return LLFE_OK;
}

// ---- host-buffer convenience entry points ----------------------------------------
// Stage through pinned memory, run on the context's stream, copy back, synchronise.
struct HostStage {
    llfe_ctx* ctx;
    uint8_t* d_in;
    uint8_t* d_out;
    uint8_t* p_in;
    uint8_t* p_out;
};

static int stage_begin(llfe_ctx* ctx, const void* h_in, size_t in_bytes, size_t out_bytes, HostStage* st) {
    const size_t a = WsCarver::need(in_bytes), b = WsCarver::need(out_bytes);
    LLFE_TRY(ensure_stage(ctx, a + b, a + b));
    st->ctx = ctx;
    st->p_in = (uint8_t*)ctx->pin;
    st->p_out = st->p_in + a;
    st->d_in = (uint8_t*)ctx->dev_stage;
    st->d_out = st->d_in + a;
    // in pieces: the DMA of a piece runs while the next one is copied into the pinned buffer
    constexpr size_t PIECE = 1 << 20;
    for (size_t off = 0; off < in_bytes; off += PIECE) {
        const size_t m = in_bytes - off < PIECE ? in_bytes - off : PIECE;
        memcpy(st->p_in + off, (const uint8_t*)h_in + off, m);
        LLFE_CUDA(cudaMemcpyAsync(st->d_in + off, st->p_in + off, m, cudaMemcpyHostToDevice, ctx->stream));
    }
    return LLFE_OK;
}

static int stage_end(HostStage* st, void* h_out, size_t out_bytes, size_t d_off) {
    LLFE_CUDA(cudaMemcpyAsync(st->p_out + d_off, st->d_out + d_off, out_bytes, cudaMemcpyDeviceToHost, st->ctx->stream));
    LLFE_CUDA(cudaStreamSynchronize(st->ctx->stream));
    memcpy(h_out, st->p_out + d_off, out_bytes);
    return LLFE_OK;
}

int llfe_shape_mask_host(llfe_ctx* ctx, const uint8_t* h_bgr, int h, int w, int low, int high, uint8_t* h_mask) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && h_bgr != nullptr && h_mask != nullptr && h > 0 && w > 0);
    const size_t p = (size_t)h * w;
    HostStage st;
    LLFE_TRY(stage_begin(ctx, h_bgr, 3 * p, p, &st));
    LLFE_TRY(llfe_shape_mask(ctx, st.d_in, 1, h, w, low, high, st.d_out));
    return stage_end(&st, h_mask, p, 0);
}

int llfe_shadow_mask_host(llfe_ctx* ctx, const uint8_t* h_bgr, int h, int w, uint8_t* h_mask, uint8_t* h_blurred,
                          uint64_t* h_sum_count) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && h_bgr != nullptr && h > 0 && w > 0 && (h_mask || h_sum_count));
    const size_t p = (size_t)h * w, pa = WsCarver::need(p);
    HostStage st;
    LLFE_TRY(stage_begin(ctx, h_bgr, 3 * p, 2 * pa + 256, &st));
    uint8_t* d_mask = st.d_out;
    uint8_t* d_blur = st.d_out + pa;
    uint64_t* d_sc = (uint64_t*)(st.d_out + 2 * pa);
    LLFE_TRY(llfe_shadow_mask(ctx, st.d_in, 1, h, w, d_mask, d_blur, d_sc));
    if (h_mask) LLFE_CUDA(cudaMemcpyAsync(st.p_out, d_mask, p, cudaMemcpyDeviceToHost, ctx->stream));
    if (h_blurred) LLFE_CUDA(cudaMemcpyAsync(st.p_out + pa, d_blur, p, cudaMemcpyDeviceToHost, ctx->stream));
    LLFE_CUDA(cudaMemcpyAsync(st.p_out + 2 * pa, d_sc, 16, cudaMemcpyDeviceToHost, ctx->stream));
    LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_mask) memcpy(h_mask, st.p_out, p);
    if (h_blurred) memcpy(h_blurred, st.p_out + pa, p);
    if (h_sum_count) memcpy(h_sum_count, st.p_out + 2 * pa, 16);
    return LLFE_OK;
}

int llfe_text_mask_host(llfe_ctx* ctx, const uint8_t* h_bgr, int h, int w, uint8_t* h_mask, int32_t* h_thresh) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && h_bgr != nullptr && h_mask != nullptr && h > 0 && w > 0);
    const size_t p = (size_t)h * w, pa = WsCarver::need(p);
    HostStage st;
    LLFE_TRY(stage_begin(ctx, h_bgr, 3 * p, pa + 256, &st));
    int32_t* d_thr = (int32_t*)(st.d_out + pa);
    LLFE_TRY(llfe_text_mask(ctx, st.d_in, 1, h, w, st.d_out, d_thr));
    LLFE_CUDA(cudaMemcpyAsync(st.p_out, st.d_out, p, cudaMemcpyDeviceToHost, ctx->stream));
    LLFE_CUDA(cudaMemcpyAsync(st.p_out + pa, d_thr, 4, cudaMemcpyDeviceToHost, ctx->stream));
    LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(h_mask, st.p_out, p);
    if (h_thresh) memcpy(h_thresh, st.p_out + pa, 4);
    return LLFE_OK;
}

int llfe_font_mask_host(llfe_ctx* ctx, const uint8_t* h_bgr, int h, int w, uint8_t* h_mask) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && h_bgr != nullptr && h_mask != nullptr && h > 0 && w > 0);
    const size_t p = (size_t)h * w;
    HostStage st;
    LLFE_TRY(stage_begin(ctx, h_bgr, 3 * p, p, &st));
    LLFE_TRY(llfe_font_mask(ctx, st.d_in, 1, h, w, st.d_out));
    return stage_end(&st, h_mask, p, 0);
}

int llfe_dominant_colors_host(llfe_ctx* ctx, const uint8_t* h_bgr, int h, int w, const int8_t* h_noise, uint64_t seed,
                              int k, int attempts, int max_iter, double eps, uint64_t rng_state, float* h_centers,
                              int32_t* h_labels, int32_t* h_n_unique, int32_t* h_k_used, double* h_compactness,
                              uint32_t* h_keys, int32_t* h_status) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && h_bgr != nullptr && h_centers != nullptr && h_labels != nullptr &&
                   h_n_unique != nullptr && h_k_used != nullptr && h > 0 && w > 0 && k >= 1);
    const size_t p = (size_t)h * w;
    const int max_unique = (int)(p < ((size_t)1 << 24) ? p : ((size_t)1 << 24));
    const size_t in_bytes = WsCarver::need(3 * p) + (h_noise ? WsCarver::need(3 * p) : 0);
    const size_t keys_b = WsCarver::need((size_t)max_unique * 4), lab_b = keys_b;
    const size_t small_b = 4096;
    // device staging: [image | noise | keys | labels | small]   pinned: [image | noise | labels | small]
    LLFE_TRY(ensure_stage(ctx, in_bytes + lab_b + small_b, in_bytes + keys_b + lab_b + small_b));
    uint8_t* pin = (uint8_t*)ctx->pin;
    uint8_t* dv = (uint8_t*)ctx->dev_stage;
    memcpy(pin, h_bgr, 3 * p);
    if (h_noise) memcpy(pin + WsCarver::need(3 * p), h_noise, 3 * p);
    LLFE_CUDA(cudaMemcpyAsync(dv, pin, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const uint8_t* d_img = dv;
    const int8_t* d_noise = h_noise ? (const int8_t*)(dv + WsCarver::need(3 * p)) : nullptr;
    uint32_t* d_keys = (uint32_t*)(dv + in_bytes);
    int32_t* d_labels = (int32_t*)(dv + in_bytes + keys_b);
    uint8_t* d_small = dv + in_bytes + keys_b + lab_b;
    int32_t* d_count = (int32_t*)d_small;             // [0..3]
    int32_t* d_kused = (int32_t*)(d_small + 8);
    uint64_t* d_rng = (uint64_t*)(d_small + 16);
    double* d_comp = (double*)(d_small + 24);
    int32_t* d_status = (int32_t*)(d_small + 32);
    float* d_centers = (float*)(d_small + 64);         // k*3 floats (k <= 32)
    uint8_t* p_small = pin + in_bytes + lab_b;
    memcpy(p_small + 16, &rng_state, 8);
    LLFE_CUDA(cudaMemcpyAsync(d_rng, p_small + 16, 8, cudaMemcpyHostToDevice, ctx->stream));
    LLFE_TRY(llfe_unique_colors(ctx, d_img, 1, h, w, d_noise, seed, 0, d_keys, nullptr, d_count, max_unique));
    LLFE_TRY(llfe_kmeans_unique(ctx, d_keys, d_count, 1, max_unique, k, attempts, max_iter, eps, d_rng, d_centers, d_labels,
                                d_comp, d_kused, nullptr, d_status));
    LLFE_CUDA(cudaMemcpyAsync(p_small, d_small, 64 + (size_t)k * 12, cudaMemcpyDeviceToHost, ctx->stream));
    LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
    int32_t n_unique = *(int32_t*)p_small;
    *h_n_unique = n_unique;
    *h_k_used = *(int32_t*)(p_small + 8);
    if (h_compactness) *h_compactness = *(double*)(p_small + 24);
    if (h_status) *h_status = *(int32_t*)(p_small + 32);
    memcpy(h_centers, p_small + 64, (size_t)k * 12);
    const size_t nl = (size_t)(n_unique < max_unique ? n_unique : max_unique);
    if (nl) {
        uint8_t* p_lab = pin + in_bytes;
        LLFE_CUDA(cudaMemcpyAsync(p_lab, d_labels, nl * 4, cudaMemcpyDeviceToHost, ctx->stream));
        LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
        memcpy(h_labels, p_lab, nl * 4);
        if (h_keys) {
            LLFE_CUDA(cudaMemcpyAsync(p_lab, d_keys, nl * 4, cudaMemcpyDeviceToHost, ctx->stream));
            LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
            memcpy(h_keys, p_lab, nl * 4);
        }
    }
    return LLFE_OK;
}

int llfe_convert_scale_abs_host(llfe_ctx* ctx, const uint8_t* h_src, size_t count, float a1, float a2, int single,
                                uint8_t* h_dst) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && h_src != nullptr && h_dst != nullptr);
    if (count == 0) return LLFE_OK;
    HostStage st;
    LLFE_TRY(stage_begin(ctx, h_src, count, count, &st));
    LLFE_TRY(llfe_convert_scale_abs(ctx, st.d_in, count, a1, a2, single, st.d_out));
    return stage_end(&st, h_dst, count, 0);
}

int llfe_gaussian_blur5_host(llfe_ctx* ctx, const uint8_t* h_src, int h, int w, int c, uint8_t* h_dst) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && h_src != nullptr && h_dst != nullptr && h > 0 && w > 0 && (c == 1 || c == 3));
    const size_t b = (size_t)h * w * c;
    HostStage st;
    LLFE_TRY(stage_begin(ctx, h_src, b, b, &st));
    LLFE_TRY(llfe_gaussian_blur5(ctx, st.d_in, 1, h, w, c, st.d_out));
    return stage_end(&st, h_dst, b, 0);
}

int llfe_png_reconstruct_host(llfe_ctx* ctx, const uint8_t* h_stream, int h, int w, int color_type, int bit_depth,
                              const uint8_t* h_palette, int palette_entries, uint8_t* h_bgr) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(h_stream != nullptr && h_bgr != nullptr && h > 0 && w > 0);
    const int64_t rb = llfe_png_rowbytes(w, color_type, bit_depth);
    LLFE_CHECK_ARG(rb > 0 && palette_entries >= 0 && palette_entries <= 256 && (color_type != 3 || h_palette != nullptr));
    const size_t in = (size_t)h * (rb + 1), out = (size_t)h * w * 3;
    // staging: the stream in; the image and, behind it, the status word out
    HostStage st;
    LLFE_TRY(stage_begin(ctx, h_stream, in, WsCarver::need(out) + 256, &st));
    uint8_t* d_pal = nullptr;
    void* ws;
    LLFE_TRY(llfe_workspace(ctx, 1024, &ws));
    if (color_type == 3) {
        uint8_t pal[768];
        memset(pal, 0, sizeof pal);
        memcpy(pal, h_palette, (size_t)palette_entries * 3);
        d_pal = (uint8_t*)ws;
        LLFE_CUDA(cudaMemcpyAsync(d_pal, pal, 768, cudaMemcpyHostToDevice, ctx->stream));   // pageable: staged before return
    }
    int32_t* d_status = (int32_t*)(st.d_out + WsCarver::need(out));
    LLFE_TRY(llfe_png_reconstruct(ctx, st.d_in, 1, h, w, color_type, bit_depth, d_pal, st.d_out, d_status));
    LLFE_TRY(stage_end(&st, h_bgr, out, 0));
    int32_t status = 0;
    LLFE_CUDA(cudaMemcpy(&status, d_status, sizeof status, cudaMemcpyDeviceToHost));
    if (status != 0) {
        llfe_set_error("llfe_png_reconstruct_host: bad adaptive filter value");
        return LLFE_E_INVALID;
    }
    return LLFE_OK;
}

int llfe_png_decode_host(llfe_ctx* ctx, const uint8_t* h_idat, size_t idat_bytes, int h, int w, int color_type, int bit_depth,
                         const uint8_t* h_palette, int palette_entries, uint8_t* h_bgr) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(h_idat != nullptr && h_bgr != nullptr && h > 0 && h <= 65535 && w > 0);
    const int64_t rb = llfe_png_rowbytes(w, color_type, bit_depth);
    LLFE_CHECK_ARG(rb > 0 && rb < 0x7fffffff && palette_entries >= 0 && palette_entries <= 256 &&
                   (color_type != 3 || h_palette != nullptr));
    const int rowbytes = (int)rb;
    const size_t stride = (size_t)rowbytes + 1, in = (size_t)h * stride, out = (size_t)h * w * 3;
    LLFE_CHECK_ARG(in < 0xffffffffull);
    const size_t a = WsCarver::need(in), b = WsCarver::need(out) + 256;
    LLFE_TRY(ensure_stage(ctx, a + b, a + b));
    uint8_t* p_in = (uint8_t*)ctx->pin;
    uint8_t* d_in = (uint8_t*)ctx->dev_stage;
    uint8_t* d_out = d_in + a;
    int32_t* d_status = (int32_t*)(d_out + WsCarver::need(out));
    const int bpp = png_filter_distance(color_type, bit_depth);
    LLFE_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int32_t), ctx->stream));
    size_t got = 0;
    int inf_rc = LLFE_OK;
    // The inflate is a serial decode on one host core and by far the longest part.  For streams worth it, it runs on a
    // helper thread and reports how far it is after every deflate block; this thread ships each band of rows to the
    // device and reconstructs it (rows above the band are finished: the wavefront kernel continues from them) while the
    // rest is still being inflated, so that only the last band's copy and kernel are left when the inflate ends.
    const int bands = in >= (size_t(1) << 20) && h >= 64 ? 8 : 1;
    if (bands > 1) {
        std::atomic<size_t> progress{0};
        std::atomic<int> done{0};
        std::thread worker([&] {
            inf_rc = llfe_inflate_zlib_progress(h_idat, idat_bytes, p_in, in, &got, &progress, ctx->opt_inflate_threads);
            done.store(1, std::memory_order_release);
        });
        int rc = LLFE_OK;
        const int rows_per_band = ceil_div(h, bands);
        for (int r0 = 0; r0 < h && rc == LLFE_OK; r0 += rows_per_band) {
            const int r1 = r0 + rows_per_band < h ? r0 + rows_per_band : h;
            const size_t need = (size_t)r1 * stride;
            while (progress.load(std::memory_order_acquire) < need && !done.load(std::memory_order_acquire)) std::this_thread::yield();
            if (progress.load(std::memory_order_acquire) < need) break;      // the stream ended early or is damaged
            const size_t off = (size_t)r0 * stride;
            cudaError_t e = cudaMemcpyAsync(d_in + off, p_in + off, need - off, cudaMemcpyHostToDevice, ctx->stream);
            if (e != cudaSuccess) {
                rc = llfe_cuda_fail(e, "cudaMemcpyAsync(band)", __FILE__, __LINE__);
                break;
            }
            rc = launch_png_unfilter_rows(ctx, d_in, 1, h, r0, r1, rowbytes, bpp, d_status);
        }
        worker.join();
        if (rc != LLFE_OK) {
            cudaStreamSynchronize(ctx->stream);
            return rc;
        }
    } else {
        inf_rc = llfe_inflate_zlib_progress(h_idat, idat_bytes, p_in, in, &got, nullptr, ctx->opt_inflate_threads);
        if (inf_rc == LLFE_OK && got == in) {
            LLFE_CUDA(cudaMemcpyAsync(d_in, p_in, in, cudaMemcpyHostToDevice, ctx->stream));
            LLFE_TRY(launch_png_unfilter_rows(ctx, d_in, 1, h, 0, h, rowbytes, bpp, d_status));
        }
    }
    if (inf_rc != LLFE_OK || got != in) {
        cudaStreamSynchronize(ctx->stream);     // bands of the valid front may be in flight
        llfe_set_error(inf_rc != LLFE_OK ? "llfe_png_decode_host: invalid or truncated deflate stream"
                                         : "llfe_png_decode_host: not enough image data");
        return LLFE_E_INVALID;
    }
    uint8_t* d_pal = nullptr;
    if (color_type == 3) {
        void* ws;
        LLFE_TRY(llfe_workspace(ctx, 1024, &ws));
        uint8_t pal[768];
        memset(pal, 0, sizeof pal);
        memcpy(pal, h_palette, (size_t)palette_entries * 3);
        d_pal = (uint8_t*)ws;
        LLFE_CUDA(cudaMemcpyAsync(d_pal, pal, 768, cudaMemcpyHostToDevice, ctx->stream));   // pageable: staged before return
    }
    LLFE_TRY(launch_png_to_bgr(ctx, d_in, 1, h, w, rowbytes, color_type, bit_depth, d_pal, d_out));
    uint8_t* p_out = p_in + a;
    LLFE_CUDA(cudaMemcpyAsync(p_out, d_out, out, cudaMemcpyDeviceToHost, ctx->stream));
    LLFE_CUDA(cudaMemcpyAsync(p_out + WsCarver::need(out), d_status, 4, cudaMemcpyDeviceToHost, ctx->stream));
    LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
    int32_t status;
    memcpy(&status, p_out + WsCarver::need(out), 4);
    if (status != 0) {
        llfe_set_error("llfe_png_decode_host: bad adaptive filter value");
        return LLFE_E_INVALID;
    }
    memcpy(h_bgr, p_out, out);
    return LLFE_OK;
}

int llfe_png_decode_adam7_host(llfe_ctx* ctx, const uint8_t* h_idat, size_t idat_bytes, int h, int w, int color_type,
                               int bit_depth, const uint8_t* h_palette, int palette_entries, uint8_t* h_bgr) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(h_idat != nullptr && h_bgr != nullptr && h > 0 && h <= 65535 && w > 0);
    const int64_t total = llfe_png_stream_bytes(w, h, color_type, bit_depth, 1);
    LLFE_CHECK_ARG(total > 0 && total < 0xffffffffll && palette_entries >= 0 && palette_entries <= 256 &&
                   (color_type != 3 || h_palette != nullptr));
    const size_t in = (size_t)total, out = (size_t)h * w * 3;
    const size_t a = WsCarver::need(in), b = WsCarver::need(out) + 256;
    LLFE_TRY(ensure_stage(ctx, a + b, a + b));
    uint8_t* p_in = (uint8_t*)ctx->pin;
    uint8_t* d_in = (uint8_t*)ctx->dev_stage;
    uint8_t* d_out = d_in + a;
    int32_t* d_status = (int32_t*)(d_out + WsCarver::need(out));
    size_t got = 0;
    if (llfe_inflate_zlib_progress(h_idat, idat_bytes, p_in, in, &got, nullptr, ctx->opt_inflate_threads) != LLFE_OK || got != in) {
        llfe_set_error("llfe_png_decode_adam7_host: invalid, truncated or short deflate stream");
        return LLFE_E_INVALID;
    }
    LLFE_CUDA(cudaMemcpyAsync(d_in, p_in, in, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* d_pal = nullptr;
    if (color_type == 3) {
        void* ws;
        LLFE_TRY(llfe_workspace(ctx, 1024, &ws));
        uint8_t pal[768];
        memset(pal, 0, sizeof pal);
        memcpy(pal, h_palette, (size_t)palette_entries * 3);
        d_pal = (uint8_t*)ws;
        LLFE_CUDA(cudaMemcpyAsync(d_pal, pal, 768, cudaMemcpyHostToDevice, ctx->stream));   // pageable: staged before return
    }
    LLFE_TRY(llfe_png_reconstruct_adam7(ctx, d_in, h, w, color_type, bit_depth, d_pal, d_out, d_status));
    uint8_t* p_out = p_in + a;
    LLFE_CUDA(cudaMemcpyAsync(p_out, d_out, out, cudaMemcpyDeviceToHost, ctx->stream));
    LLFE_CUDA(cudaMemcpyAsync(p_out + WsCarver::need(out), d_status, 4, cudaMemcpyDeviceToHost, ctx->stream));
    LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
    int32_t status;
    memcpy(&status, p_out + WsCarver::need(out), 4);
    if (status != 0) {
        llfe_set_error("llfe_png_decode_adam7_host: bad adaptive filter value");
        return LLFE_E_INVALID;
    }
    memcpy(h_bgr, p_out, out);
    return LLFE_OK;
}

int llfe_jpeg_decode_host(llfe_ctx* ctx, const uint8_t* h_buf, size_t len, int h, int w, uint8_t* h_bgr) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(h_buf != nullptr && h_bgr != nullptr && h > 0 && w > 0 && len >= 4);
    const size_t cap = llfe_jpeg_stage_bytes(h, w);
    LLFE_TRY(ensure_stage(ctx, cap, cap));
    return llfe_jpeg_decode_impl(ctx, h_buf, len, h, w, h_bgr, (uint8_t*)ctx->pin, (uint8_t*)ctx->dev_stage, cap);
}

int llfe_pil_resize_lanczos_host(llfe_ctx* ctx, const uint8_t* h_src, int sh, int sw, int c, int fx, int fy,
                                 const int32_t* reduce_box, const float* box, uint8_t* h_dst, int dh, int dw) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(h_src != nullptr && h_dst != nullptr && box != nullptr && sh > 0 && sw > 0 && dh > 0 && dw > 0);
    LLFE_CHECK_ARG((c == 1 || c == 3) && fx >= 1 && fy >= 1 && ((fx == 1 && fy == 1) || reduce_box != nullptr));
    const bool reduce = fx > 1 || fy > 1;
    int rh = sh, rw = sw;
    if (reduce) {
        LLFE_CHECK_ARG(reduce_box[0] >= 0 && reduce_box[1] >= 0 && reduce_box[2] > reduce_box[0] &&
                       reduce_box[3] > reduce_box[1] && reduce_box[2] <= sw && reduce_box[3] <= sh);
        rw = ceil_div(reduce_box[2] - reduce_box[0], fx);
        rh = ceil_div(reduce_box[3] - reduce_box[1], fy);
    }
    const size_t in = (size_t)sh * sw * c, mid = reduce ? WsCarver::need((size_t)rh * rw * c) : 0, out = (size_t)dh * dw * c;
    HostStage st;
    LLFE_TRY(stage_begin(ctx, h_src, in, mid + out, &st));
    const uint8_t* src = st.d_in;
    if (reduce) {
        LLFE_TRY(llfe_pil_reduce(ctx, st.d_in, 1, sh, sw, c, reduce_box, fx, fy, st.d_out));
        src = st.d_out;
    }
    LLFE_TRY(llfe_pil_resample_lanczos(ctx, src, 1, rh, rw, c, box, st.d_out + mid, dh, dw));
    return stage_end(&st, h_dst, out, mid);
}

int llfe_resize_area_host(llfe_ctx* ctx, const uint8_t* h_src, int sh, int sw, int c, uint8_t* h_dst, int dh, int dw) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && h_src != nullptr && h_dst != nullptr && sh > 0 && sw > 0 && dh > 0 && dw > 0);
    const size_t in = (size_t)sh * sw * c, out = (size_t)dh * dw * c;
    HostStage st;
    LLFE_TRY(stage_begin(ctx, h_src, in, out, &st));
    LLFE_TRY(llfe_resize_area(ctx, st.d_in, 1, sh, sw, c, st.d_out, dh, dw));
    return stage_end(&st, h_dst, out, 0);
}

int llfe_resize_linear_host(llfe_ctx* ctx, const uint8_t* h_src, int sh, int sw, int c, uint8_t* h_dst, int dh, int dw) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && h_src != nullptr && h_dst != nullptr && sh > 0 && sw > 0 && dh > 0 && dw > 0);
    const size_t in = (size_t)sh * sw * c, out = (size_t)dh * dw * c;
    HostStage st;
    LLFE_TRY(stage_begin(ctx, h_src, in, out, &st));
    LLFE_TRY(llfe_resize_linear(ctx, st.d_in, 1, sh, sw, c, st.d_out, dh, dw));
    return stage_end(&st, h_dst, out, 0);
}

int llfe_resize_lanczos4_host(llfe_ctx* ctx, const uint8_t* h_src, int sh, int sw, int c, uint8_t* h_dst, int dh, int dw) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(ctx != nullptr && h_src != nullptr && h_dst != nullptr && sh > 0 && sw > 0 && dh > 0 && dw > 0);
    const size_t in = (size_t)sh * sw * c, out = (size_t)dh * dw * c;
    HostStage st;
    LLFE_TRY(stage_begin(ctx, h_src, in, out, &st));
    LLFE_TRY(llfe_resize_lanczos4(ctx, st.d_in, 1, sh, sw, c, st.d_out, dh, dw));
    return stage_end(&st, h_dst, out, 0);
}

// ---- contours (k_contours.cu) on host buffers: only headers + polygon vertices come back ----------------------------
static int contours_fetch(llfe_ctx* ctx, HostStage* st, size_t hdr_off, size_t pts_off, size_t cnt_off, int max_contours,
                          int max_points, int32_t* h_headers, int32_t* h_points, int32_t* h_counts) {
    LLFE_CUDA(cudaMemcpyAsync(st->p_out + cnt_off, st->d_out + cnt_off, 16, cudaMemcpyDeviceToHost, ctx->stream));
    LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
    int32_t cnt[4];
    memcpy(cnt, st->p_out + cnt_off, 16);
    const size_t nh = (size_t)(cnt[0] < max_contours ? cnt[0] : max_contours) * 40;
    const size_t np = (size_t)(cnt[1] < max_points ? cnt[1] : max_points) * 8;
    if (nh) LLFE_CUDA(cudaMemcpyAsync(st->p_out + hdr_off, st->d_out + hdr_off, nh, cudaMemcpyDeviceToHost, ctx->stream));
    if (np) LLFE_CUDA(cudaMemcpyAsync(st->p_out + pts_off, st->d_out + pts_off, np, cudaMemcpyDeviceToHost, ctx->stream));
    if (nh || np) LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
    if (nh) memcpy(h_headers, st->p_out + hdr_off, nh);
    if (np) memcpy(h_points, st->p_out + pts_off, np);
    memcpy(h_counts, cnt, 16);
    return LLFE_OK;
}

int llfe_contours_external_host(llfe_ctx* ctx, const uint8_t* h_mask, int h, int w, int64_t min_area2, int32_t* h_headers,
                                int max_contours, int32_t* h_points, int max_points, int32_t* h_counts) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(h_mask != nullptr && h_headers != nullptr && h_counts != nullptr && h > 0 && w > 0 && max_contours > 0 &&
                   max_points >= 0 && (h_points != nullptr || max_points == 0));
    const size_t p = (size_t)h * w, hb = WsCarver::need((size_t)max_contours * 40), pb = WsCarver::need((size_t)max_points * 8);
    HostStage st;
    LLFE_TRY(stage_begin(ctx, h_mask, p, hb + pb + 256, &st));
    LLFE_TRY(llfe_contours_external(ctx, st.d_in, 1, h, w, min_area2, (int32_t*)st.d_out, max_contours,
                                    (int32_t*)(st.d_out + hb), max_points, (int32_t*)(st.d_out + hb + pb)));
    return contours_fetch(ctx, &st, 0, hb, hb + pb, max_contours, max_points, h_headers, h_points, h_counts);
}

int llfe_shape_contours_host(llfe_ctx* ctx, const uint8_t* h_bgr, int h, int w, int low, int high, int64_t min_area2,
                             uint8_t* h_mask, int32_t* h_headers, int max_contours, int32_t* h_points, int max_points,
                             int32_t* h_counts) {
    LLFE_ENTER(ctx);
    LLFE_CHECK_ARG(h_bgr != nullptr && h_headers != nullptr && h_counts != nullptr && h > 0 && w > 0 && max_contours > 0 &&
                   max_points >= 0 && (h_points != nullptr || max_points == 0));
    const size_t p = (size_t)h * w, ma = WsCarver::need(p), hb = WsCarver::need((size_t)max_contours * 40),
                 pb = WsCarver::need((size_t)max_points * 8);
    HostStage st;
    LLFE_TRY(stage_begin(ctx, h_bgr, 3 * p, ma + hb + pb + 256, &st));
    uint8_t* d_mask = st.d_out;
    LLFE_TRY(llfe_shape_mask(ctx, st.d_in, 1, h, w, low, high, d_mask));
    LLFE_TRY(llfe_contours_external(ctx, d_mask, 1, h, w, min_area2, (int32_t*)(st.d_out + ma), max_contours,
                                    (int32_t*)(st.d_out + ma + hb), max_points, (int32_t*)(st.d_out + ma + hb + pb)));
    if (h_mask) LLFE_CUDA(cudaMemcpyAsync(st.p_out, d_mask, p, cudaMemcpyDeviceToHost, ctx->stream));
    LLFE_TRY(contours_fetch(ctx, &st, ma, ma + hb, ma + hb + pb, max_contours, max_points, h_headers, h_points, h_counts));
    if (h_mask) memcpy(h_mask, st.p_out, p);
    return LLFE_OK;
}

}  // extern "C"

