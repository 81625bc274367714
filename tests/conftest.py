import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # OpenCV 4.13's multi-threaded filters are not deterministic on tiny images (GaussianBlur of an
    # 11x11x3 image returns a partly wrong second row in ~1 of 3 fresh processes with 8 threads;
    # single-threaded it is stable and equals the oracle).  The checker must be deterministic.
    try:
        import cv2

        cv2.setNumThreads(1)
    except Exception:
        pass


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference (see oracle/make_golden.py)."""
    gdir = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(gdir, "golden.json")) as f:
        meta = json.load(f)
    npz = np.load(os.path.join(gdir, "golden.npz"))
    arrays = {k.replace("__", "/"): npz[k] for k in npz.files}
    return meta, arrays


@pytest.fixture(scope="session")
def golden_inputs(golden):
    """Rebuild the seeded inputs and prove they are the bytes the fixtures were made from."""
    import hashlib
    from low_level_feature_extraction_b200.synth import design_image, noise_image

    meta, _ = golden
    imgs = {
        "design_270x480_s1": design_image(270, 480, 1),
        "design_360x640_s2": design_image(360, 640, 2),
        "noise_96x160_s3": noise_image(96, 160, 3),
        "design_101x203_s4": design_image(101, 203, 4),
    }
    for k, im in imgs.items():
        assert hashlib.sha256(im.tobytes()).hexdigest() == meta["cases"][k]["input_sha256"], k
    return imgs


@pytest.fixture(scope="session")
def llfe():
    """The ctypes binding of libllfe.so (GPU tests call through the C ABI)."""
    import low_level_feature_extraction_b200 as pkg
    return pkg
