"""The `/extract-colors`, `/extract-shapes` and `/extract-shadows` routers the reference's README lists
(/root/reference/README.md:221-226; their modules are gone from the tree, only the services survive as stale
byte-code), rebuilt on the drop-in services: same preprocessing entry (`validate_and_preprocess_image`,
app/services/analyze/utils.py:90-152, any failure -> HTTP 400 as at :147-152), same JSON shapes (`ColorFeatures`,
app/api/v1/models/analyze.py:157-204; the `analyze_shapes` dict, shape_analyzer pyc L183-189; the shadow level
string, shadow_analyzer pyc L12-31).

SURVEY 8(f)4's batching queue: when the app is created with a `RequestBatcher`, concurrent requests of one image
shape share one staged batched launch (colours + shape mask + contours + shadow mask in a single `llfe_analyze`
pass per batch) instead of one C call sequence per request; every request still receives exactly what the
one-at-a-time services return for its image (tests/test_gpu_api.py).

    uvicorn low_level_feature_extraction_b200.services.api:app      # LLFE_BATCH=0 disables the batcher

This file is wiring only (control plane): no arithmetic lives here.
"""
from __future__ import annotations

import contextlib
import inspect
import os
import time
from typing import Any, Dict, Optional

from fastapi import APIRouter, FastAPI, File, HTTPException, Query, UploadFile, status
from starlette.concurrency import run_in_threadpool

from .color_extractor import ColorExtractor
from .shadow_analyzer import ShadowAnalyzer
from .shape_analyzer import ShapeAnalyzer
from .utils import validate_and_preprocess_image

PREPROCESSING_MODES = ("auto", "none", "high_quality", "performance")   # utils.py:118-143


def create_app(batcher: Optional[Any] = None, default_preprocessing: str = "auto") -> FastAPI:
    """batcher: a `services.batching.RequestBatcher` (or None: every request runs the single-image services)."""
    @contextlib.asynccontextmanager
    async def lifespan(_app):
        yield
        if batcher is not None:
            batcher.close()

    app = FastAPI(title="Low-level feature extraction (B200 path)", lifespan=lifespan)
    router = APIRouter(responses={status.HTTP_400_BAD_REQUEST: {"description": "Invalid request"},
                                  status.HTTP_500_INTERNAL_SERVER_ERROR: {"description": "Internal server error"}})

    async def load(file: UploadFile, preprocessing: str):
        if preprocessing not in PREPROCESSING_MODES:
            raise HTTPException(status_code=status.HTTP_400_BAD_REQUEST, detail=f"unknown preprocessing mode {preprocessing!r}")
        data = await file.read()
        request_id = f"req-{int(time.time())}"                       # analyze.py:78
        return await validate_and_preprocess_image(data, request_id, preprocessing)   # raises HTTPException(400)

    async def features(image) -> Optional[Dict[str, Any]]:
        """One batched pass for everything, when a batcher with the default palette size is attached."""
        if batcher is None:
            return None
        return await run_in_threadpool(batcher.analyze, image)

    def guard(fn):
        async def wrapped(*a, **k):
            try:
                return await fn(*a, **k)
            except HTTPException:
                raise
            except Exception as e:                                      # analyze.py:133-139
                raise HTTPException(status_code=status.HTTP_500_INTERNAL_SERVER_ERROR,
                                    detail=f"An unexpected error occurred: {e}")
        wrapped.__name__ = fn.__name__
        wrapped.__doc__ = fn.__doc__
        wrapped.__signature__ = inspect.signature(fn)
        return wrapped

    @app.get("/")
    async def root():
        return {"service": "low-level-feature-extraction", "endpoints": ["/extract-colors", "/extract-shapes", "/extract-shadows"],
                "batched": batcher is not None}

    @router.post("/extract-colors")
    @guard
    async def extract_colors(file: UploadFile = File(...), n_colors: int = Query(5, ge=1, le=16),
                             preprocessing: str = Query(default_preprocessing)):
        image = await load(file, preprocessing)
        if batcher is not None and n_colors == getattr(batcher, "n_colors", None):
            colors = (await features(image))["colors"]
        else:
            colors = await run_in_threadpool(ColorExtractor.extract_colors, image, n_colors)
        return colors.model_dump() if hasattr(colors, "model_dump") else colors.dict()

    @router.post("/extract-shapes")
    @guard
    async def extract_shapes(file: UploadFile = File(...), preprocessing: str = Query(default_preprocessing)):
        image = await load(file, preprocessing)
        if batcher is not None and getattr(batcher, "shapes", False):
            return (await features(image))["shapes"]
        return await run_in_threadpool(ShapeAnalyzer.analyze_shapes, image)

    @router.post("/extract-shadows")
    @guard
    async def extract_shadows(file: UploadFile = File(...), preprocessing: str = Query(default_preprocessing)):
        image = await load(file, preprocessing)
        if batcher is not None:
            level = (await features(image))["shadow_level"]
        else:
            level = await run_in_threadpool(ShadowAnalyzer.analyze_shadow_level, image)
        return {"shadow_level": level}

    app.include_router(router)
    app.state.batcher = batcher
    return app


def _default_app() -> FastAPI:
    if os.environ.get("LLFE_BATCH", "1") != "0":
        from .batching import RequestBatcher

        return create_app(RequestBatcher(device=int(os.environ.get("LLFE_DEVICE", "0"))))
    return create_app(None)


def __getattr__(name):   # `uvicorn ...api:app` builds the app (and its GPU context) only when asked for
    if name == "app":
        a = _default_app()
        globals()["app"] = a
        return a
    raise AttributeError(name)
