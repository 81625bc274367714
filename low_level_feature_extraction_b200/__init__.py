"""B200-native image hot path of Kira7dn/Low_Level_Feature_Extraction.

Preprocessing (resize, grayscale, Gaussian blur, contrast scaling), dominant
colour palette (noise + unique colours + k-means) and the Canny / adaptive /
Otsu masks behind shape and shadow analysis, as hand-written sm_100a CUDA
kernels behind a C ABI (`libllfe.so`, include/llfe.h).

    from low_level_feature_extraction_b200 import engine          # torch-tensor batch API
    from low_level_feature_extraction_b200.services import ...     # drop-in service classes

There is no CPU fallback: importing the package loads libllfe.so and fails if
it has not been built (`python low_level_feature_extraction_b200/build.py`).
"""
from ._native import Context, LlfeError, load_library, PROTOTYPES, LIB_PATH  # noqa: F401

load_library()  # fail loudly at import if the native library is missing


def engine(device=None):
    from .ops import engine as _engine

    return _engine(device)


__version__ = "0.1.0"
