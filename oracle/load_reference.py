"""Import the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE.  `/root/reference` does not exist on the GPU box, so
nothing that runs there may call this; it exists for `make_golden.py` and the
CPU-only cross-checks in tests/test_reference_live.py (skipped when the tree
is absent).  The stale byte-code services are loaded as-is with
SourcelessFileLoader (CPython 3.12 magic matches) -- SURVEY.md section 8(c).
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys

REF_ROOT = os.environ.get("LLFE_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "app", "services"))


def _load_pyc(name: str, path: str):
    loader = importlib.machinery.SourcelessFileLoader(name, path)
    spec = importlib.util.spec_from_loader(name, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def load():
    """Returns a dict of the reference's classes / functions on the path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    pyc = os.path.join(REF_ROOT, "app", "services", "__pycache__")
    from app.services.analyze.color_extractor import ColorExtractor
    from app.services.analyze.image_processor import ImageProcessor
    from app.services.analyze.utils import validate_and_preprocess_image
    import app.services.analyze.image_processor as ip

    sys.modules.setdefault("app.services.image_processor", ip)  # old module path used by the stale pyc
    shape = _load_pyc("ref_shape_analyzer", os.path.join(pyc, "shape_analyzer.cpython-312.pyc"))
    shadow = _load_pyc("ref_shadow_analyzer", os.path.join(pyc, "shadow_analyzer.cpython-312.pyc"))
    transformer = _load_pyc("ref_image_transformer", os.path.join(pyc, "image_transformer.cpython-312.pyc"))
    return {
        "ColorExtractor": ColorExtractor,
        "ImageProcessor": ImageProcessor,
        "validate_and_preprocess_image": validate_and_preprocess_image,
        "ShapeAnalyzer": shape.ShapeAnalyzer,
        "ShadowAnalyzer": shadow.ShadowAnalyzer,
        "ImageTransformer": transformer.ImageTransformer,
    }
