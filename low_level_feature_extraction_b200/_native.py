"""ctypes binding of libllfe.so (the C ABI declared in include/llfe.h).

There is no CPU fallback: if the shared library is missing this module raises
at import time, and if no CUDA device is usable `Context()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libllfe.so")

LLFE_OK = 0
LLFE_E_INVALID = -1
LLFE_E_UNSUPPORTED = -4
ERRORS = {-1: "LLFE_E_INVALID", -2: "LLFE_E_CUDA", -3: "LLFE_E_NOMEM", -4: "LLFE_E_UNSUPPORTED", -5: "LLFE_E_NODEVICE"}


class LlfeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


vp, i32, u64, sz, f32, f64 = C.c_void_p, C.c_int, C.c_uint64, C.c_size_t, C.c_float, C.c_double

# name -> (restype, argtypes); every symbol include/llfe.h declares
PROTOTYPES = {
    "llfe_version": (i32, []),
    "llfe_last_error": (C.c_char_p, []),
    "llfe_device_count": (i32, []),
    "llfe_create": (i32, [i32, C.POINTER(vp)]),
    "llfe_destroy": (i32, [vp]),
    "llfe_set_stream": (i32, [vp, vp]),
    "llfe_use_own_stream": (i32, [vp]),
    "llfe_sync": (i32, [vp]),
    "llfe_set_option": (i32, [vp, C.c_char_p, C.c_int64]),
    "llfe_set_debug_buffer": (i32, [vp, C.c_char_p, vp, sz]),
    "llfe_launch_count": (u64, [vp]),
    "llfe_sm_count": (i32, [vp]),
    "llfe_profile_begin": (i32, [vp]),
    "llfe_profile_end": (i32, [vp, C.c_char_p, sz]),
    "llfe_malloc": (i32, [vp, sz, C.POINTER(vp)]),
    "llfe_free": (i32, [vp, vp]),
    "llfe_malloc_host": (i32, [vp, sz, C.POINTER(vp)]),
    "llfe_free_host": (i32, [vp, vp]),
    "llfe_memcpy_h2d": (i32, [vp, vp, vp, sz]),
    "llfe_memcpy_d2h": (i32, [vp, vp, vp, sz]),
    "llfe_memset": (i32, [vp, vp, i32, sz]),
    "llfe_bgr2gray": (i32, [vp, vp, i32, i32, i32, vp]),
    "llfe_bgr2rgb": (i32, [vp, vp, i32, i32, i32, vp]),
    "llfe_gaussian_blur5": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "llfe_gray_blur5": (i32, [vp, vp, i32, i32, i32, vp]),
    "llfe_canny": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "llfe_hysteresis": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "llfe_dilate3": (i32, [vp, vp, i32, i32, i32, vp]),
    "llfe_shape_mask": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "llfe_adaptive_threshold": (i32, [vp, vp, i32, i32, i32, i32, vp, vp]),
    "llfe_shadow_mask": (i32, [vp, vp, i32, i32, i32, vp, vp, vp]),
    "llfe_font_mask": (i32, [vp, vp, i32, i32, i32, vp]),
    "llfe_otsu": (i32, [vp, vp, i32, i32, i32, i32, vp, vp]),
    "llfe_text_mask": (i32, [vp, vp, i32, i32, i32, vp, vp]),
    "llfe_resize_area": (i32, [vp, vp, i32, i32, i32, i32, vp, i32, i32]),
    "llfe_resize_linear": (i32, [vp, vp, i32, i32, i32, i32, vp, i32, i32]),
    "llfe_resize_lanczos4": (i32, [vp, vp, i32, i32, i32, i32, vp, i32, i32]),
    "llfe_convert_scale_abs": (i32, [vp, vp, sz, f32, f32, i32, vp]),
    "llfe_unique_colors": (i32, [vp, vp, i32, i32, i32, vp, u64, i32, vp, vp, vp, i32]),
    "llfe_kmeans_unique": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, f64, vp, vp, vp, vp, vp, vp, vp]),
    "llfe_kmeans_lloyd": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, f64, i32, vp, vp, vp, vp, vp]),
    "llfe_kmeans_pixels_step": (i32, [vp, vp, sz, i32, vp, vp, vp, vp]),
    "llfe_kmeans_pixels_zero": (i32, [vp, i32, vp, vp]),
    "llfe_kmeans_pixels_farthest": (i32, [vp, vp, sz, i32, vp, i32, vp, C.c_uint32, vp, i32, C.c_uint32, vp]),
    "llfe_kmeans_hist_farthest": (i32, [vp, vp, sz, i32, vp, i32, vp, vp]),
    "llfe_kmeans_update": (i32, [vp, i32, vp, vp, i32, f64, vp, vp, vp, i32]),
    "llfe_p2p_mailbox_bytes": (sz, []),
    "llfe_ipc_export": (i32, [vp, vp, vp]),
    "llfe_ipc_open": (i32, [vp, vp, C.POINTER(vp)]),
    "llfe_ipc_close": (i32, [vp, vp]),
    "llfe_kmeans_update_p2p": (i32, [vp, i32, vp, vp, i32, i32, vp, i32, f64, vp, vp, vp]),
    "llfe_pixels_histogram": (i32, [vp, vp, sz, vp]),
    "llfe_histogram_compact": (i32, [vp, vp, i32, i32, vp, vp, sz, vp, i32]),
    "llfe_kmeans_hist_step": (i32, [vp, vp, vp, sz, i32, vp, vp, vp, vp, vp]),
    "llfe_p2p_barrier": (i32, [vp, vp, i32, i32]),
    "llfe_histogram_pull_reduce": (i32, [vp, vp, i32, i32, vp]),
    "llfe_kmeans_hist_lloyd": (i32, [vp, vp, vp, sz, vp, i32, vp, vp, vp, vp, i32, i32, i32, f64, vp, vp, vp, i32]),
    "llfe_hist_labels_to_lut": (i32, [vp, vp, vp, sz, vp]),
    "llfe_pixels_lookup": (i32, [vp, vp, sz, vp, vp]),
    "llfe_contours_external": (i32, [vp, vp, i32, i32, i32, C.c_int64, vp, i32, vp, i32, vp]),
    "llfe_contours_external_host": (i32, [vp, vp, i32, i32, C.c_int64, vp, i32, vp, i32, vp]),
    "llfe_shape_contours_host": (i32, [vp, vp, i32, i32, i32, i32, C.c_int64, vp, vp, i32, vp, i32, vp]),
    "llfe_pipeline": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, u64, vp, vp, i32]),
    "llfe_analyze": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, u64, vp, vp, i32, i32, i32, i32, f64, vp, vp, vp, vp,
                           vp, vp]),
    "llfe_pack_mask_bits": (i32, [vp, vp, i32, i32, i32, vp]),
    "llfe_mask_bits_words_per_row": (i32, [i32]),
    "llfe_expand_mask_bits_host": (i32, [vp, i32, i32, i32, vp, i32]),
    "llfe_shape_mask_host": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "llfe_shadow_mask_host": (i32, [vp, vp, i32, i32, vp, vp, vp]),
    "llfe_text_mask_host": (i32, [vp, vp, i32, i32, vp, vp]),
    "llfe_font_mask_host": (i32, [vp, vp, i32, i32, vp]),
    "llfe_resize_area_host": (i32, [vp, vp, i32, i32, i32, vp, i32, i32]),
    "llfe_resize_linear_host": (i32, [vp, vp, i32, i32, i32, vp, i32, i32]),
    "llfe_resize_lanczos4_host": (i32, [vp, vp, i32, i32, i32, vp, i32, i32]),
    "llfe_gaussian_blur5_host": (i32, [vp, vp, i32, i32, i32, vp]),
    "llfe_png_rowbytes": (C.c_int64, [i32, i32, i32]),
    "llfe_png_reconstruct": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "llfe_png_reconstruct_host": (i32, [vp, vp, i32, i32, i32, i32, vp, i32, vp]),
    "llfe_inflate_zlib": (i32, [vp, sz, vp, sz, C.POINTER(sz)]),
    "llfe_inflate_zlib_mt": (i32, [vp, sz, vp, sz, C.POINTER(sz), i32]),
    "llfe_png_decode_host": (i32, [vp, vp, sz, i32, i32, i32, i32, vp, i32, vp]),
    "llfe_png_stream_bytes": (C.c_int64, [i32, i32, i32, i32, i32]),
    "llfe_png_reconstruct_adam7": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, vp]),
    "llfe_png_decode_adam7_host": (i32, [vp, vp, sz, i32, i32, i32, i32, vp, i32, vp]),
    "llfe_jpeg_info": (i32, [vp, sz, vp]),
    "llfe_jpeg_coefficients": (i32, [vp, sz, vp, sz, C.POINTER(sz)]),
    "llfe_jpeg_decode_host": (i32, [vp, vp, sz, i32, i32, vp]),
    "llfe_pil_reduce": (i32, [vp, vp, i32, i32, i32, i32, vp, i32, i32, vp]),
    "llfe_pil_resample_lanczos": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, i32, i32]),
    "llfe_pil_resize_lanczos_host": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, i32, i32]),
    "llfe_convert_scale_abs_host": (i32, [vp, vp, sz, f32, f32, i32, vp]),
    "llfe_dominant_colors_host": (i32, [vp, vp, i32, i32, vp, u64, i32, i32, i32, f64, u64, vp, vp, vp, vp, vp, vp, vp]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library() -> C.CDLL:
    """dlopen libllfe.so and declare every prototype.  Does not touch CUDA."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `python low_level_feature_extraction_b200/build.py` "
                    "(there is no CPU fallback)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(lib, name)  # AttributeError here = header/library mismatch
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def _ptr(x) -> int | None:
    """Device/host address of a torch tensor, numpy array, int, or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(f"cannot take the address of {type(x)}")


class Context:
    """One llfe_ctx: a device, a stream, a workspace arena."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = vp()
        rc = self.lib.llfe_create(device, C.byref(h))
        if rc != LLFE_OK:
            raise LlfeError(rc, self.lib.llfe_last_error().decode())
        self.handle = h
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.lib.llfe_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, name: str, *args):
        fn = getattr(self.lib, name)
        rc = fn(self.handle, *[a if isinstance(a, (int, float, bytes)) else _ptr(a) for a in args])
        if rc != LLFE_OK:
            raise LlfeError(rc, self.lib.llfe_last_error().decode())

    def set_stream(self, cuda_stream: int | None):
        """cuda_stream: a cudaStream_t as int; 0 / None = the legacy default stream."""
        self.call("llfe_set_stream", cuda_stream or None)

    def use_own_stream(self):
        self.call("llfe_use_own_stream")

    def sync(self):
        self.call("llfe_sync")

    def set_option(self, name: str, value: int):
        """Parity-test path toggles: "unfused", "hyst_strips" (include/llfe.h)."""
        self.call("llfe_set_option", name.encode(), int(value))

    def set_debug_buffer(self, name: str, tensor=None):
        """Phase-clock records of the "kmeans" / "hysteresis" kernels into a CUDA tensor (None = off)."""
        if tensor is None:
            self.call("llfe_set_debug_buffer", name.encode(), None, 0)
        else:
            self.call("llfe_set_debug_buffer", name.encode(), tensor, tensor.numel() * tensor.element_size())

    def profile_begin(self):
        self.call("llfe_profile_begin")

    def profile_end(self) -> dict:
        import json

        buf = C.create_string_buffer(1 << 14)
        rc = self.lib.llfe_profile_end(self.handle, buf, len(buf))
        if rc != LLFE_OK:
            raise LlfeError(rc, self.lib.llfe_last_error().decode())
        return json.loads(buf.value.decode())

    @property
    def launches(self) -> int:
        return int(self.lib.llfe_launch_count(self.handle))

    @property
    def sm_count(self) -> int:
        return int(self.lib.llfe_sm_count(self.handle))
