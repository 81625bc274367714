// Entry points that are declared in llfe.h but not implemented yet.
#include "llfe_common.cuh"
#define TODO(name) llfe_set_error(name ": not implemented yet"); return LLFE_E_UNSUPPORTED
extern "C" {
int llfe_unique_colors(llfe_ctx*, const uint8_t*, int, int, int, const int8_t*, uint64_t, uint32_t*, uint32_t*, int32_t*, int) { TODO("llfe_unique_colors"); }
int llfe_kmeans_unique(llfe_ctx*, const uint32_t*, const int32_t*, int, int, int, int, int, double, const uint64_t*, float*, int32_t*, double*, int32_t*) { TODO("llfe_kmeans_unique"); }
int llfe_kmeans_lloyd(llfe_ctx*, const uint32_t*, const uint32_t*, const int32_t*, int, int, int, int, double, int, const float*, float*, int32_t*, int32_t*, uint64_t*) { TODO("llfe_kmeans_lloyd"); }
int llfe_kmeans_pixels_step(llfe_ctx*, const uint8_t*, size_t, int, const float*, uint64_t*, uint8_t*) { TODO("llfe_kmeans_pixels_step"); }
int llfe_kmeans_update(llfe_ctx*, int, const uint64_t*, float*, int, double, int32_t*, double*) { TODO("llfe_kmeans_update"); }
int llfe_pipeline(llfe_ctx*, const uint8_t*, int, int, int, int, int, uint8_t*, uint8_t*, uint64_t*, const int8_t*, uint64_t, uint32_t*, int32_t*, int) { TODO("llfe_pipeline"); }
}
