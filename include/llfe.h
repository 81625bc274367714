/*
 * llfe.h -- C ABI of libllfe.so: the B200 (sm_100a) image hot path of
 * Kira7dn/Low_Level_Feature_Extraction.
 *
 * The reference has no FFI for this path: every arithmetic step is a call from
 * a Python service method into OpenCV/NumPy.  Each entry point below replaces
 * one such call site (cited as file:line relative to the reference tree; "pyc"
 * = app/services/__pycache__/<module>.cpython-312.pyc, line = original source
 * line recorded in the code object).  INTEGRATION.md shows the ctypes stubs a
 * maintainer of the reference would add to bind them.
 *
 * Conventions
 *   - plain C types only; no CUDA or torch types in the signatures.
 *   - every function returns 0 on success or a negative LLFE_E_* code;
 *     llfe_last_error() returns a thread-local human-readable message.
 *   - images are uint8, HWC, BGR channel order (the OpenCV convention of the
 *     reference), C-contiguous; a batch is `n` images of identical h x w stored
 *     back to back.
 *   - pointers named d_* are DEVICE pointers, pointers named h_* are HOST
 *     pointers.  All d_* work is enqueued on the context's stream and is
 *     asynchronous; *_host entry points copy in, run and copy out, and return
 *     after the result is in host memory.
 *   - the caller owns every buffer; the context owns only its workspace arena
 *     and pinned staging buffers.
 *   - one context per (process, device); calls on one context are serialised on
 *     its stream (the reference handles one request at a time per worker).  Every
 *     entry point makes the context's device current for the duration of the call and
 *     restores the caller's device on return, so a context may be used from any thread
 *     (a request-handler pool) whatever device that thread has current.
 *   - there is NO CPU fallback: if no CUDA device is usable, llfe_create fails.
 */
#ifndef LLFE_H
#define LLFE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLFE_OK 0
#define LLFE_E_INVALID (-1)  /* bad argument */
#define LLFE_E_CUDA (-2)     /* CUDA runtime error (message has the detail) */
#define LLFE_E_NOMEM (-3)    /* allocation failed */
#define LLFE_E_UNSUPPORTED (-4)
#define LLFE_E_NODEVICE (-5)

typedef struct llfe_ctx llfe_ctx;

/* ---- library / context ------------------------------------------------ */
int llfe_version(void);
const char* llfe_last_error(void);
int llfe_device_count(void);
int llfe_create(int device, llfe_ctx** out);
int llfe_destroy(llfe_ctx* ctx);
/* Use an existing CUDA stream (a cudaStream_t passed as void*), e.g. torch's
 * current stream.  NULL is a valid stream: the CUDA legacy default stream (what
 * torch calls its default stream).  llfe_use_own_stream goes back to the
 * non-blocking stream the context created for itself. */
int llfe_set_stream(llfe_ctx* ctx, void* cuda_stream);
int llfe_use_own_stream(llfe_ctx* ctx);
int llfe_sync(llfe_ctx* ctx);
/* Path toggles for parity tests (production leaves both 0): "unfused" = 1 runs the per-stage kernels
 * instead of the fused front kernel, "hyst_strips" = 1 the multi-launch strip hysteresis instead of the
 * cluster kernel, "shadow_inline" = 1 the adaptive threshold inside the front kernel instead of k_shadow,
 * "serial" = 1 one stream for the two chains of llfe_pipeline / llfe_analyze, "contour_segments" = 0 follows every
 * border of llfe_contours_external with one thread (default 1: calls on one or two images cut long borders into
 * segments that are followed in parallel), "contour_cut_shift" = log2 of the rows / columns between the cuts
 * (default 6; tests use 0..3), "chunk" = images per front-kernel launch, "inflate_threads" = decoders per deflate stream
 * in llfe_png_decode_host / llfe_png_decode_adam7_host (1..8, default 4; see llfe_inflate_zlib_mt).  Unknown names fail
 * with LLFE_E_INVALID. */
int llfe_set_option(llfe_ctx* ctx, const char* name, int64_t value);
/* Register a device buffer the "kmeans" / "hysteresis" kernels write per-CTA phase clocks to (tools/debug/);
 * d_buf = NULL switches the records off.  The pointer is validated with cudaPointerGetAttributes (device
 * memory of this context's device) and the kernels write only when `bytes` covers the launch. */
int llfe_set_debug_buffer(llfe_ctx* ctx, const char* name, void* d_buf, size_t bytes);
/* Number of kernels this context has launched so far (for accounting). */
uint64_t llfe_launch_count(llfe_ctx* ctx);
int llfe_sm_count(llfe_ctx* ctx);
/* Per-kernel timing with CUDA events on the context's stream.  Between begin and
 * end every kernel launch is bracketed by an event pair; end synchronises and
 * writes {"kernel name": {"ms": total, "launches": count}, ...} into json. */
int llfe_profile_begin(llfe_ctx* ctx);
int llfe_profile_end(llfe_ctx* ctx, char* json, size_t cap);

/* ---- memory helpers (so a C / ctypes host needs no CUDA runtime) -------- */
int llfe_malloc(llfe_ctx* ctx, size_t bytes, void** d_out);
int llfe_free(llfe_ctx* ctx, void* d_ptr);
int llfe_malloc_host(llfe_ctx* ctx, size_t bytes, void** h_out); /* pinned */
int llfe_free_host(llfe_ctx* ctx, void* h_ptr);
int llfe_memcpy_h2d(llfe_ctx* ctx, void* d_dst, const void* h_src, size_t bytes); /* async on stream */
int llfe_memcpy_d2h(llfe_ctx* ctx, void* h_dst, const void* d_src, size_t bytes); /* async on stream */
int llfe_memset(llfe_ctx* ctx, void* d_dst, int value, size_t bytes);

/* ---- per-op entry points (device pointers, batched) ---------------------- */

/* cv2.cvtColor(img, COLOR_BGR2GRAY): shape_analyzer pyc L18, shadow_analyzer
 * pyc L8, app/services/analyze/text_extractor.py:27, font_detector.py:28.
 * Y = (3735 B + 19235 G + 9798 R + 16384) >> 15. */
int llfe_bgr2gray(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, uint8_t* d_gray);

/* cv2.cvtColor(img, COLOR_BGR2RGB): app/services/analyze/color_extractor.py:151. */
int llfe_bgr2rgb(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, uint8_t* d_rgb);

/* cv2.GaussianBlur(src, (5,5), 0) on u8 with c = 1 or 3 channels:
 * shape_analyzer pyc L21, shadow_analyzer pyc L9, image_transformer pyc L103. */
int llfe_gaussian_blur5(llfe_ctx* ctx, const uint8_t* d_src, int n, int h, int w, int c, uint8_t* d_dst);

/* gray + blur fused: ShadowAnalyzer.preprocess_image, shadow_analyzer pyc L5-10. */
int llfe_gray_blur5(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, uint8_t* d_blurred);

/* cv2.Canny(gray, low, high) (aperture 3, L1 norm): shape_analyzer pyc L24. */
int llfe_canny(llfe_ctx* ctx, const uint8_t* d_gray, int n, int h, int w, int low, int high, uint8_t* d_edges);

/* The hysteresis stage of cv2.Canny on its own (shape_analyzer pyc L24, SURVEY.md A.3
 * step 5): d_weak / d_strong are u8 maps (non-zero = pixel kept by the non-maximum
 * suppression / kept and above the high threshold; strong pixels outside the weak set
 * are ignored).
 * d_edges = 255 for every weak pixel whose 8-connected component of weak pixels contains
 * a strong pixel, else 0; dilate != 0 folds the 3x3 dilate of pyc L27-28 in. */
int llfe_hysteresis(llfe_ctx* ctx, const uint8_t* d_weak, const uint8_t* d_strong, int n, int h, int w, int dilate,
                    uint8_t* d_edges);

/* cv2.dilate(src, ones(3,3), iterations=1): shape_analyzer pyc L27-28. */
int llfe_dilate3(llfe_ctx* ctx, const uint8_t* d_src, int n, int h, int w, uint8_t* d_dst);

/* ShapeAnalyzer.preprocess_image (shape_analyzer pyc L6-30), fused:
 * gray -> blur5 -> Canny(low, high) -> dilate3.  d_mask is (n,h,w) u8 in {0,255}. */
int llfe_shape_mask(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, int low, int high, uint8_t* d_mask);

/* cv2.adaptiveThreshold(src, 255, ADAPTIVE_THRESH_GAUSSIAN_C, THRESH_BINARY_INV,
 * 11, C): shadow_analyzer pyc L17-18, font_detector.py:31-35.  If d_sum_count is
 * not NULL it receives, per image, {sum of src where mask==255, count of
 * mask==255} as two uint64 (the masked mean of shadow_analyzer pyc L21-24). */
int llfe_adaptive_threshold(llfe_ctx* ctx, const uint8_t* d_gray, int n, int h, int w, int C, uint8_t* d_mask,
                            uint64_t* d_sum_count);

/* ShadowAnalyzer.analyze_shadow_level up to the scalar (shadow_analyzer pyc
 * L12-24), fused: gray -> blur5 -> adaptive(11, 2) -> masked sum/count.
 * d_blurred may be NULL when the caller does not need the blurred image. */
int llfe_shadow_mask(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, uint8_t* d_mask, uint8_t* d_blurred,
                     uint64_t* d_sum_count);

/* FontDetector.preprocess_image (font_detector.py:16-37): gray -> adaptive(11,2). */
int llfe_font_mask(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, uint8_t* d_mask);

/* cv2.threshold(gray, 0, 255, THRESH_BINARY + THRESH_OTSU): text_extractor.py:40.
 * d_thresh receives the Otsu level per image (int32).  If invert_if_light != 0 the
 * mask is inverted when mean(mask) > 127 (text_extractor.py:43-44). */
int llfe_otsu(llfe_ctx* ctx, const uint8_t* d_gray, int n, int h, int w, int invert_if_light, uint8_t* d_mask,
              int32_t* d_thresh);

/* TextExtractor.preprocess_image for h >= 30 and w >= 100 (text_extractor.py:15-46):
 * gray -> Otsu -> invert if mostly white. */
int llfe_text_mask(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, uint8_t* d_mask, int32_t* d_thresh);

/* cv2.resize(src, (dw, dh), interpolation=INTER_AREA), down-scaling, c = 1 or 3:
 * app/services/analyze/utils.py:127, image_processor.py:112, image_transformer
 * pyc L53, L170-173. */
int llfe_resize_area(llfe_ctx* ctx, const uint8_t* d_src, int n, int sh, int sw, int c, uint8_t* d_dst, int dh,
                     int dw);

/* cv2.resize(src, (dw, dh), interpolation=INTER_LINEAR) on u8, c = 1 or 3 (OpenCV's 11-bit
 * fixed-point path; up- and down-scaling): the `performance` preprocessing mode,
 * app/services/analyze/utils.py:136-143. */
int llfe_resize_linear(llfe_ctx* ctx, const uint8_t* d_src, int n, int sh, int sw, int c, uint8_t* d_dst, int dh,
                       int dw);

/* cv2.resize(src, (dw, dh), interpolation=INTER_LANCZOS4) on u8, c = 1 or 3 (8-tap, OpenCV's 11-bit
 * fixed-point path): the `high_quality` preprocessing mode, app/services/analyze/utils.py:128-135. */
int llfe_resize_lanczos4(llfe_ctx* ctx, const uint8_t* d_src, int n, int sh, int sw, int c, uint8_t* d_dst, int dh,
                         int dw);

/* ---- image decode and Pillow thumbnailing ---------------------------------------------------------------
 * cv2.imdecode(buf, IMREAD_COLOR) of a non-interlaced PNG (app/services/analyze/utils.py:108-109,
 * image_processor.py:62-66, :208-211) after the host has inflated the IDAT stream: d_stream holds, per image,
 * h scanlines of [filter type][llfe_png_rowbytes bytes] and is reconstructed IN PLACE (PNG filters None / Sub / Up /
 * Average / Paeth), then converted to BGR u8 the way OpenCV configures libpng (16-bit samples -> high byte, gray 1/2/4
 * bits scaled to 0..255, palette looked up in d_palette = n x 256 x 3 RGB bytes zero-padded, alpha / tRNS dropped).
 * color_type / bit_depth are IHDR's.  d_status[i] != 0: image i has an invalid filter byte (libpng fails the decode). */
int64_t llfe_png_rowbytes(int w, int color_type, int bit_depth);
int llfe_png_reconstruct(llfe_ctx* ctx, uint8_t* d_stream, int n, int h, int w, int color_type, int bit_depth,
                         const uint8_t* d_palette, uint8_t* d_bgr, int32_t* d_status);

/* Adam7-interlaced PNGs (IHDR interlace method 1), one image per call: the stream holds seven reduced images, each
 * filtered on its own; llfe_png_stream_bytes = size of the scanline stream for either interlace method (host-only). */
int64_t llfe_png_stream_bytes(int w, int h, int color_type, int bit_depth, int interlace);
int llfe_png_reconstruct_adam7(llfe_ctx* ctx, uint8_t* d_stream, int h, int w, int color_type, int bit_depth,
                               const uint8_t* d_palette, uint8_t* d_bgr, int32_t* d_status);

/* Host-only: inflate a zlib stream (RFC 1950 / 1951; the concatenated IDAT payloads of a PNG) into out, at most out_cap
 * bytes.  Failure (LLFE_E_INVALID) wherever zlib's inflate fails -- invalid code sets or symbols, distance too far back,
 * truncated input, Adler-32 mismatch; input beyond the point where the output is full is ignored, as libpng does once the
 * image is complete.  No context, no device. */
int llfe_inflate_zlib(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len);
/* The same result from `threads` (1..8) decoders working on the one stream: all but the first start at a dynamic-block
 * header they find in the middle of the compressed data and decode into symbols that may refer to the 32 KB window they
 * do not know; when the decoder in front arrives at exactly that bit, at a block boundary, the symbols become bytes.  A
 * start nobody arrives at, or a failing worker, only costs time (the decoder in front goes on alone); streams with fewer
 * than 128 KB of compressed data per decoder use fewer.  llfe_png_decode_host uses this with
 * llfe_set_option("inflate_threads") decoders (default 4). */
int llfe_inflate_zlib_mt(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len, int threads);

/* Host-only: is `buf` a JFIF JPEG of the subset the device path decodes (8-bit, Huffman; baseline with one interleaved
 * scan or progressive; gray or YCbCr 4:4:4 / 4:2:2 / 4:2:0; no Exif orientation / Adobe marker)?  out[0] = width, out[1] = height.
 * LLFE_E_UNSUPPORTED = a JPEG outside the subset (the caller uses cv2.imdecode), LLFE_E_INVALID = damaged header.
 * llfe_jpeg_coefficients (host-only, for tests): the entropy-decoded quantised coefficients, blocks [by][bx][64] in natural
 * order, component after component. */
int llfe_jpeg_info(const uint8_t* buf, size_t len, int32_t* out);
int llfe_jpeg_coefficients(const uint8_t* buf, size_t len, int16_t* out, size_t cap, size_t* count);

/* Pillow's ImagingReduce(im, (fx, fy), box) and ImagingResample(im, (dw, dh), LANCZOS, box) on u8, c = 1 or 3: the two
 * steps of `pil_image.thumbnail(size, Image.Resampling.LANCZOS)` (image_processor.py:221-224; reducing_gap = 2.0).
 * box = host int32[4] / float[4] (left, upper, right, lower) in source pixels.  llfe_pil_reduce writes
 * ceil((box[3]-box[1])/fy) x ceil((box[2]-box[0])/fx) pixels. */
int llfe_pil_reduce(llfe_ctx* ctx, const uint8_t* d_src, int n, int sh, int sw, int c, const int32_t* box, int fx, int fy,
                    uint8_t* d_dst);
int llfe_pil_resample_lanczos(llfe_ctx* ctx, const uint8_t* d_src, int n, int sh, int sw, int c, const float* box,
                              uint8_t* d_dst, int dh, int dw);

/* cv2.convertScaleAbs(x, alpha=a1, beta=0) followed by (alpha=a2, beta=0), the
 * pair of calls of ImageTransformer.adjust_brightness_contrast (image_transformer
 * pyc L139-142) fused into one pass.  Pass a2 = 1.0f with single = 1 for one call. */
int llfe_convert_scale_abs(llfe_ctx* ctx, const uint8_t* d_src, size_t count, float a1, float a2, int single,
                           uint8_t* d_dst);

/* ---- palette (ColorExtractor) -------------------------------------------- */

/* Noise + unique colours: color_extractor.py:151 (BGR2RGB), :224-225 (noise add +
 * clip) and :177 (np.unique(axis=0)).  d_noise is the reference's int8 noise
 * tensor (n,h,w,3) in RGB order, or NULL to generate noise of the same
 * distribution on the device from `seed` (throughput mode; not bit-equal to
 * NumPy's MT19937 stream).  d_keys receives, per image, the sorted unique colours
 * as R<<16|G<<8|B (== np.unique row order) at d_keys + i*max_unique; d_count[i]
 * receives the number of unique colours (values above max_unique mean truncated: only the
 * first max_unique keys were written; come back with a longer list, see llfe_kmeans_unique).
 * first_image: index of image 0 of this call in the caller's numbering -- the device noise is a
 * pure function of (seed, first_image + i, pixel position), so a single image of a larger batch
 * can be redone later with the same noise.
 * If d_hist is not NULL it receives the pixel count of each unique colour
 * (uint32, same layout as d_keys) -- the weights of the per-pixel k-means mode. */
int llfe_unique_colors(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, const int8_t* d_noise,
                       uint64_t seed, int first_image, uint32_t* d_keys, uint32_t* d_hist, int32_t* d_count,
                       int max_unique);

/* cv2.kmeans(float32(unique), K, None, (EPS+MAX_ITER, max_iter, eps), attempts,
 * KMEANS_PP_CENTERS) on each image's unique-colour list: color_extractor.py:189-197.
 * rng_state[i] is the cv::RNG state the reference would start image i with
 * (cv2.setRNGSeed(s) => s, 0 => 0xffffffff).  Outputs per image: centers (k x 3
 * float32, RGB), labels (int32 per unique colour), compactness (double), and
 * k_used = min(k, n_unique), and cluster_sizes (k int32: unique colours per cluster,
 * the np.bincount(labels) of color_extractor.py:232).  Optional outputs may be NULL.
 * max_iter is clamped to [2, 100] as cv::kmeans does with criteria.maxCount (the reference's 200 is 100).
 * k = 1 or an image with fewer than 2 unique colours: no clustering, k_used = min(1, n_unique), centre 0 = the
 * first unique colour, labels = 0 (color_extractor.py:185-186 returns the unique list itself in that case; the
 * Python layer rebuilds it from d_keys).
 * A list that was truncated (d_count[i] > max_unique) is NOT clustered: k_used[i] = -1 and status bit
 * LLFE_KMEANS_TRUNCATED tell the caller to redo that image with a list of d_count[i] entries.
 * d_status (optional, int32 per image): LLFE_KMEANS_LONG_SUMS = some cluster's channel sum reached 2^24, where
 * cv2's sequential float32 centre sums round; that regime is reproduced operation by operation (k_kmeans_seq.cuh),
 * the bit is informational. */
#define LLFE_KMEANS_LONG_SUMS 1
#define LLFE_KMEANS_TRUNCATED 2
int llfe_kmeans_unique(llfe_ctx* ctx, const uint32_t* d_keys, const int32_t* d_count, int n, int max_unique, int k,
                       int attempts, int max_iter, double eps, const uint64_t* d_rng_state, float* d_centers,
                       int32_t* d_labels, double* d_compactness, int32_t* d_k_used, int32_t* d_cluster_sizes,
                       int32_t* d_status);

/* Lloyd iterations from given initial centres (the "seeded mode" of SURVEY.md
 * A.8) over weighted colours (d_weights NULL = all ones).  exact_sums = 0:
 * cv2's float32 centre rule (sum * (1.f/count)); exact_sums = 1: the pinned
 * per-pixel rule c = float(double(sum)/double(count)).  d_iters gets the
 * iteration count per problem. */
int llfe_kmeans_lloyd(llfe_ctx* ctx, const uint32_t* d_keys, const uint32_t* d_weights, const int32_t* d_count, int n,
                      int max_unique, int k, int max_iter, double eps, int exact_sums, const float* d_init_centers,
                      float* d_centers, int32_t* d_labels, int32_t* d_iters, uint64_t* d_sums_counts);

/* One per-pixel k-means step for a row shard of ONE image (multi-GPU mode,
 * SURVEY.md 8(e)): assign every pixel of d_bgr (rows x w) to the nearest of k
 * centres (float32 RGB) and ADD the exact per-cluster sums into
 * d_sums_counts[k][4] = {sum R, sum G, sum B, count} (uint64).  The caller
 * zeroes the accumulator, all-reduces it across ranks (ncclSum) and calls
 * llfe_kmeans_update.  d_state_or_null is the state array of llfe_kmeans_update: when it
 * says "converged" (state[1]) or "frozen" (state[3]) the call is a no-op, so a host can
 * enqueue several iterations back to back and look at the state once per batch. */
int llfe_kmeans_pixels_step(llfe_ctx* ctx, const uint8_t* d_bgr, size_t n_pixels, int k, const float* d_centers,
                            uint64_t* d_sums_counts, uint8_t* d_labels_or_null, const int32_t* d_state_or_null);

/* Zero the k x 4 accumulator for the next iteration unless the state says converged / frozen. */
int llfe_kmeans_pixels_zero(llfe_ctx* ctx, int k, uint64_t* d_sums_counts, const int32_t* d_state_or_null);

/* Empty-cluster repair support for the per-pixel mode: among the pixels of this
 * shard whose nearest centre (under d_centers) is `donor`, find the one farthest
 * (float32 distance) from h_base3 = the donor's provisional mean; ties go to the
 * highest pixel index.  *d_out = max(*d_out, ((dist bits << 32) | (index_base +
 * local pixel index)) + 1); the caller zeroes it first and all-reduces with max.
 * h_skip[n_skip] (n_skip <= 32) lists global pixel indices to ignore: the pixels
 * earlier repairs of the same update already moved out of the donor.  With
 * want_dist_bits != 0 only pixels at exactly that float32 distance from h_base3 are
 * considered: when llfe_kmeans_hist_farthest has already found the answer's distance, the
 * pass costs little more than reading the rows (0 = consider every pixel). */
int llfe_kmeans_pixels_farthest(llfe_ctx* ctx, const uint8_t* d_bgr, size_t n_pixels, int k, const float* d_centers,
                                int donor, const float* h_base3, uint32_t index_base, const uint32_t* h_skip, int n_skip,
                                uint32_t want_dist_bits, uint64_t* d_out);

/* The same search over (key, count) entries, distance only: *d_out_bits = max(*d_out_bits, bits of the
 * largest float32 distance to h_base3 among the colours assigned to `donor`) (all-reduce with max). */
int llfe_kmeans_hist_farthest(llfe_ctx* ctx, const uint32_t* d_keys, size_t n, int k, const float* d_centers, int donor,
                              const float* h_base3, uint32_t* d_out_bits);

/* Centre update + convergence test from (all-reduced) sums: c =
 * float(double(sum)/double(count)); d_state (4 x int32): [0] = iteration counter (in/out),
 * [1] = converged flag (out), [2] = number of empty clusters (out), [3] = frozen (out: set
 * together with [2] != 0; nothing else is updated -- the host repairs the sums, clears
 * [2] and [3], and calls again).  A converged or frozen state makes the call a no-op.
 * d_shift receives max_k |c - old|^2 (double).  The first call (iteration 0) never
 * reports convergence, as in cv2's KMEANS_USE_INITIAL_LABELS mode.
 * d_consumed_or_null (k x 4), when given, receives the sums this update used; with
 * zero_sums != 0 (needs d_consumed) d_sums_counts is cleared afterwards, so a per-rank
 * accumulator can be all-reduced in place every iteration and is empty again for the next
 * assignment -- the per-iteration loop is then step, all-reduce, update, nothing else. */
int llfe_kmeans_update(llfe_ctx* ctx, int k, uint64_t* d_sums_counts, float* d_centers, int max_iter, double eps,
                       int32_t* d_state, double* d_shift, uint64_t* d_consumed_or_null, int zero_sums);

/* ---- fused all-reduce + centre update over peer memory (config 5 on N > 1 GPUs of one node) --------------------------
 * Replaces the pair {ncclAllReduce of the k x 4 sums, llfe_kmeans_update} of the per-iteration loop ("per
 * iteration each rank produces K x 3 channel sums + K counts as uint64 => all-reduce, then every rank recomputes identical
 * centres") by ONE kernel per rank that exchanges the partial sums through NVLink peer stores.  Set-up, once per process
 * group: every rank allocates a mailbox of llfe_p2p_mailbox_bytes() with llfe_malloc, zeroes it, exports it
 * (llfe_ipc_export -> 64 opaque bytes, the cudaIpcMemHandle), exchanges the handles out of band, opens the peers'
 * (llfe_ipc_open) and keeps the `world` device pointers (its own at index `rank`) in a device array; a barrier before the
 * first use.  Every rank then calls llfe_kmeans_update_p2p the same number of times: d_partial_sums (this rank's k x 4
 * accumulator) is consumed and cleared, d_totals receives the global sums, centres / state / shift behave exactly as in
 * llfe_kmeans_update (a converged or frozen state makes the call a no-op on every rank alike). */
size_t llfe_p2p_mailbox_bytes(void);
int llfe_ipc_export(llfe_ctx* ctx, void* d_ptr, uint8_t* handle64);
int llfe_ipc_open(llfe_ctx* ctx, const uint8_t* handle64, void** d_peer_out);
int llfe_ipc_close(llfe_ctx* ctx, void* d_peer);
int llfe_kmeans_update_p2p(llfe_ctx* ctx, int k, uint64_t* d_partial_sums, void* const* d_mailboxes, int rank, int world,
                           float* d_centers, int max_iter, double eps, int32_t* d_state, double* d_shift, uint64_t* d_totals);

/* ---- colour-histogram form of the per-pixel k-means (BASELINE config 5) ------
 * A u8 image has at most 2^24 distinct colours and the centre update needs only exact
 * integer sums, so the rows are streamed ONCE into a count table and the Lloyd iterations
 * run over the distinct colours weighted by their pixel counts: the same labels, sums and
 * centres as llfe_kmeans_pixels_step (color_extractor.py:189-196 applied per pixel), without
 * re-reading the image every iteration. */

/* d_hist[key] += number of pixels of colour key = (R << 16) + (G << 8) + B, 2^24 uint32
 * bins (the caller zeroes the table once and all-reduces it across ranks with ncclSum;
 * the image as a whole must have fewer than 2^31 pixels so that no bin overflows). */
int llfe_pixels_histogram(llfe_ctx* ctx, const uint8_t* d_bgr, size_t n_pixels, uint32_t* d_hist);

/* Ordered compaction of the non-empty bins into (key, count) entries, key ascending =
 * np.unique's (R, G, B) order.  The table is cut into blocks of 2048 keys; only blocks b with
 * b % parts == part are emitted (rank r of a world of G passes part = r, parts = G).
 * *d_n = number of entries of that part; at most `cap` entries are written (cap = 0 with NULL
 * arrays just counts).  packed != 0: d_hist holds ONLY this part's blocks, back to back (block b of
 * the table at slot b / parts; 8192 % parts == 0) -- what a reduce-scatter over the block-transposed
 * table leaves on each rank, so that only 1/G of the table crosses NVLink instead of all of it. */
int llfe_histogram_compact(llfe_ctx* ctx, const uint32_t* d_hist, int part, int parts, uint32_t* d_keys_or_null,
                           uint32_t* d_counts_or_null, size_t cap, int32_t* d_n, int packed);

/* llfe_kmeans_pixels_step over (key, count) entries: nearest centre per distinct colour
 * (cv2's float32 distance, first minimum), d_sums_counts += count * {R, G, B, 1}; optional
 * one label byte per entry; same d_state_or_null convention.  d_n_or_null: the number of entries as
 * llfe_histogram_compact left it on the device (min(n, *d_n) entries are used), so that the host never
 * has to read the count back between the compaction and the iterations. */
int llfe_kmeans_hist_step(llfe_ctx* ctx, const uint32_t* d_keys, const uint32_t* d_counts, size_t n, int k,
                          const float* d_centers, uint64_t* d_sums_counts, uint8_t* d_labels_or_null,
                          const int32_t* d_state_or_null, const int32_t* d_n_or_null);

/* The colour table across the ranks of one node without NCCL (config 5): every rank's 2^24-bin table lives in memory
 * exported to the peers (llfe_ipc_export); llfe_p2p_barrier (device-side barrier through the mailboxes: all tables are
 * complete) and then llfe_histogram_pull_reduce, which reads this rank's interleaved 2048-key blocks out of every table in
 * d_tables[world] over NVLink, sums them and writes the packed share (2^24 / world bins) that llfe_histogram_compact takes
 * with packed = 1.  world must divide 8192. */
int llfe_p2p_barrier(llfe_ctx* ctx, void* const* d_mailboxes, int rank, int world);
int llfe_histogram_pull_reduce(llfe_ctx* ctx, const uint32_t* const* d_tables, int rank, int world, uint32_t* d_share);

/* The whole per-iteration loop {llfe_kmeans_hist_step, exchange of the sums, llfe_kmeans_update} as ONE persistent
 * cooperative kernel per rank: up to `iterations` Lloyd iterations over this rank's (key, count) entries, with the sums
 * exchanged through the peers' mailboxes (d_mailboxes as for llfe_kmeans_update_p2p; NULL with world = 1) and the centres,
 * state, shift and totals updated as by llfe_kmeans_update.  The kernel stops early when the state says converged or
 * frozen (the host repairs an empty cluster and calls again).  Every rank calls it with the same `iterations`. */
int llfe_kmeans_hist_lloyd(llfe_ctx* ctx, const uint32_t* d_keys, const uint32_t* d_counts, size_t n, const int32_t* d_n_or_null,
                           int k, float* d_centers, uint64_t* d_partial_sums, uint8_t* d_labels_or_null,
                           void* const* d_mailboxes_or_null, int rank, int world, int max_iter, double eps, int32_t* d_state,
                           double* d_shift, uint64_t* d_totals, int iterations);
/* d_lut[key] = label for every entry (d_lut: 2^24 bytes, zeroed by the caller; ranks
 * all-reduce it with ncclSum since every colour belongs to exactly one part). */
int llfe_hist_labels_to_lut(llfe_ctx* ctx, const uint32_t* d_keys, const uint8_t* d_labels, size_t n, uint8_t* d_lut);

/* d_labels[p] = d_lut[colour of pixel p]: the per-pixel labels of the last assignment. */
int llfe_pixels_lookup(llfe_ctx* ctx, const uint8_t* d_bgr, size_t n_pixels, const uint8_t* d_lut, uint8_t* d_labels);

/* ---- fused service pipelines ---------------------------------------------- */

/* colours + shapes + shadows from one read of the image (BASELINE config 4):
 * shape mask (llfe_shape_mask), shadow mask + sum/count (llfe_shadow_mask) and
 * the unique-colour list (llfe_unique_colors) of every image. Any output group
 * may be disabled by passing NULL for its mask / keys pointer. */
int llfe_pipeline(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, int low, int high, uint8_t* d_shape_mask,
                  uint8_t* d_shadow_mask, uint64_t* d_shadow_sum_count, const int8_t* d_noise, uint64_t seed,
                  uint32_t* d_keys, int32_t* d_count, int max_unique);

/* llfe_pipeline + llfe_kmeans_unique in ONE call (what BatchAnalyzer runs per batch): the mask chain (front kernel,
 * shadow kernel, hysteresis) and the colour chain (colour pass, compaction, k-means) are independent, so the call
 * enqueues them on two streams of the context -- forked from and joined back into the context's stream, i.e. to
 * the caller it behaves like any other asynchronous call on that stream -- and the latency-bound kernels of one
 * chain fill issue slots the other leaves idle.  Results are the same as the two separate calls
 * (llfe_set_option(ctx, "serial", 1) keeps everything on one stream).  k-means arguments as llfe_kmeans_unique;
 * d_keys / d_count are required (the lists the palette is computed from), d_labels may be NULL. */
int llfe_analyze(llfe_ctx* ctx, const uint8_t* d_bgr, int n, int h, int w, int low, int high, uint8_t* d_shape_mask,
                 uint8_t* d_shadow_mask, uint64_t* d_shadow_sum_count, const int8_t* d_noise, uint64_t seed, uint32_t* d_keys,
                 int32_t* d_count, int max_unique, int k, int attempts, int max_iter, double eps, const uint64_t* d_rng_state,
                 float* d_centers, int32_t* d_labels, int32_t* d_k_used, int32_t* d_cluster_sizes, int32_t* d_status);

/* ---- external contours of a mask (cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)) -----------------------
 * Replaces the call in ShapeAnalyzer.extract_shapes / analyze_shapes (app/services/__pycache__/shape_analyzer.cpython-312.pyc,
 * source lines 76 and 140) and FontDetector.detect_text_regions (app/services/analyze/font_detector.py:51-55).
 * Any non-zero mask byte is foreground, as in OpenCV.  Per image the call writes
 *   d_counts  (n, 4) int32: {external contours found, points written, 1 if max_points was too small, internal};
 *   d_headers (n, max_contours) records of 40 bytes = int32 {start (y * w + x of the contour's first point), npts, offset
 *             (index of its first point in the image's point array, -1 if its points were not written), min x, min y,
 *             max x, max y, 0} + int64 {2 * signed area (Green's formula: |value| / 2 == cv2.contourArea)};
 *   d_points  (n, max_points, 2) int32 (x, y): the CHAIN_APPROX_SIMPLE vertices, in OpenCV's order, of the contours with
 *             |2 * area| >= min_area2 (the reference drops contours with cv2.contourArea < 100: min_area2 = 200).
 * The records are in no particular order; cv2 returns contours by DESCENDING `start`.  A caller that sees
 * counts[0] > max_contours or counts[2] != 0 repeats the call with larger buffers.  max_points = 0 (d_points may be NULL)
 * returns headers only (bounding boxes, areas). */
int llfe_contours_external(llfe_ctx* ctx, const uint8_t* d_mask, int n, int h, int w, int64_t min_area2, int32_t* d_headers,
                           int max_contours, int32_t* d_points, int max_points, int32_t* d_counts);

/* ---- bit-packed masks for the host path -----------------------------------------------
 * The masks the reference returns (u8, {0, 255}) carry one bit per pixel.  For host consumers the library can pack a
 * device mask into a bit plane (bit x & 31 of word x >> 5, rows padded to llfe_mask_bits_words_per_row(w) 32-bit words,
 * images back to back), so that P/8 instead of P bytes cross PCIe, and expand a plane that has arrived in host memory
 * into the u8 array with `threads` host threads (no CUDA involved: plain pointers, any host memory). */
int llfe_pack_mask_bits(llfe_ctx* ctx, const uint8_t* d_mask, int n, int h, int w, uint32_t* d_bits);
int llfe_mask_bits_words_per_row(int w);
int llfe_expand_mask_bits_host(const uint32_t* h_bits, int n, int h, int w, uint8_t* h_mask, int threads);

/* ---- host-buffer convenience entry points (single image, synchronous) ------ */
int llfe_shape_mask_host(llfe_ctx* ctx, const uint8_t* h_bgr, int h, int w, int low, int high, uint8_t* h_mask);
/* contours of a host mask / of the shape mask of a host image (ShapeAnalyzer: gray -> blur -> Canny -> dilate -> contours in
 * one call; h_mask may be NULL when the caller does not need the mask itself) */
int llfe_contours_external_host(llfe_ctx* ctx, const uint8_t* h_mask, int h, int w, int64_t min_area2, int32_t* h_headers,
                                int max_contours, int32_t* h_points, int max_points, int32_t* h_counts);
int llfe_shape_contours_host(llfe_ctx* ctx, const uint8_t* h_bgr, int h, int w, int low, int high, int64_t min_area2,
                             uint8_t* h_mask, int32_t* h_headers, int max_contours, int32_t* h_points, int max_points,
                             int32_t* h_counts);
int llfe_shadow_mask_host(llfe_ctx* ctx, const uint8_t* h_bgr, int h, int w, uint8_t* h_mask, uint8_t* h_blurred,
                          uint64_t* h_sum_count);
int llfe_text_mask_host(llfe_ctx* ctx, const uint8_t* h_bgr, int h, int w, uint8_t* h_mask, int32_t* h_thresh);
int llfe_font_mask_host(llfe_ctx* ctx, const uint8_t* h_bgr, int h, int w, uint8_t* h_mask);
int llfe_resize_area_host(llfe_ctx* ctx, const uint8_t* h_src, int sh, int sw, int c, uint8_t* h_dst, int dh, int dw);
int llfe_resize_linear_host(llfe_ctx* ctx, const uint8_t* h_src, int sh, int sw, int c, uint8_t* h_dst, int dh, int dw);
int llfe_resize_lanczos4_host(llfe_ctx* ctx, const uint8_t* h_src, int sh, int sw, int c, uint8_t* h_dst, int dh, int dw);
/* llfe_png_reconstruct on a host stream (h_palette: palette_entries x 3 RGB bytes or NULL); returns LLFE_E_INVALID with
 * the message "bad adaptive filter value" for a stream libpng would reject. */
int llfe_png_reconstruct_host(llfe_ctx* ctx, const uint8_t* h_stream, int h, int w, int color_type, int bit_depth,
                              const uint8_t* h_palette, int palette_entries, uint8_t* h_bgr);
/* The whole decode of one non-interlaced PNG after the chunk walk: h_idat = the concatenated IDAT payloads; inflated
 * straight into pinned memory (llfe_inflate_zlib's decoder, on a helper thread for streams of 1 MB and more so that the
 * rows already inflated are copied and reconstructed meanwhile), reconstructed and converted on the device.  A damaged
 * stream (short, invalid, bad filter byte) returns LLFE_E_INVALID. */
int llfe_png_decode_host(llfe_ctx* ctx, const uint8_t* h_idat, size_t idat_bytes, int h, int w, int color_type, int bit_depth,
                         const uint8_t* h_palette, int palette_entries, uint8_t* h_bgr);
/* llfe_png_decode_host for an Adam7-interlaced file (no overlap of inflate and reconstruction: the passes interleave rows) */
int llfe_png_decode_adam7_host(llfe_ctx* ctx, const uint8_t* h_idat, size_t idat_bytes, int h, int w, int color_type,
                               int bit_depth, const uint8_t* h_palette, int palette_entries, uint8_t* h_bgr);
/* cv2.imdecode(buf, IMREAD_COLOR) of a baseline or progressive JPEG (utils.py:108-109, image_processor.py:62-66, :208-211): Huffman
 * decoding on the calling thread into pinned memory, libjpeg-turbo's islow IDCT, fancy chroma up-sampling and YCbCr -> BGR
 * conversion on the device, bit for bit.  h, w from llfe_jpeg_info.  Files outside the subset / damaged data:
 * LLFE_E_UNSUPPORTED / LLFE_E_INVALID (the caller uses cv2.imdecode). */
int llfe_jpeg_decode_host(llfe_ctx* ctx, const uint8_t* h_buf, size_t len, int h, int w, uint8_t* h_bgr);
/* Image.resize(size, LANCZOS, box, reducing_gap) below PIL/Image.py's Python layer on a host image: an optional
 * ImagingReduce by (fx, fy) over reduce_box (fx = fy = 1: none) followed by ImagingResample with `box` (in pixels of the
 * reduced image) to dh x dw; the intermediate stays on the device. */
int llfe_pil_resize_lanczos_host(llfe_ctx* ctx, const uint8_t* h_src, int sh, int sw, int c, int fx, int fy,
                                 const int32_t* reduce_box, const float* box, uint8_t* h_dst, int dh, int dw);
int llfe_gaussian_blur5_host(llfe_ctx* ctx, const uint8_t* h_src, int h, int w, int c, uint8_t* h_dst);
int llfe_convert_scale_abs_host(llfe_ctx* ctx, const uint8_t* h_src, size_t count, float a1, float a2, int single,
                                uint8_t* h_dst);
/* ColorExtractor._get_dominant_colors on one host image (color_extractor.py:151,
 * :224-225, :173-201): BGR2RGB + noise + np.unique + cv2.kmeans.  h_noise: the int8
 * noise tensor (h,w,3, RGB order) or NULL for device noise from `seed`.  h_centers:
 * k*3 floats; h_labels: room for min(h*w, 2^24) int32 (n_unique are written).  The unique-colour list is
 * sized from h*w, so it is never truncated.  h_keys (optional, same room as h_labels) receives the sorted
 * unique colours R<<16|G<<8|B -- what :185-186 returns as "centres" when fewer than two clusters are asked
 * for or possible; h_status (optional) the LLFE_KMEANS_* bits of llfe_kmeans_unique. */
int llfe_dominant_colors_host(llfe_ctx* ctx, const uint8_t* h_bgr, int h, int w, const int8_t* h_noise, uint64_t seed,
                              int k, int attempts, int max_iter, double eps, uint64_t rng_state, float* h_centers,
                              int32_t* h_labels, int32_t* h_n_unique, int32_t* h_k_used, double* h_compactness,
                              uint32_t* h_keys, int32_t* h_status);

#ifdef __cplusplus
}
#endif
#endif /* LLFE_H */
