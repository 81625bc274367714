"""The routers on the GPU: every endpoint returns what the one-at-a-time drop-in service returns for the uploaded image,
with and without the batching queue; concurrent uploads share launches."""
import threading

import cv2
import numpy as np
import pytest

torch = pytest.importorskip("torch")
fastapi = pytest.importorskip("fastapi")
pytestmark = pytest.mark.gpu

from fastapi.testclient import TestClient  # noqa: E402

from oracle import refpath  # noqa: E402
from low_level_feature_extraction_b200.services import ShadowAnalyzer, ShapeAnalyzer  # noqa: E402
from low_level_feature_extraction_b200.services.api import create_app  # noqa: E402
from low_level_feature_extraction_b200.services.batching import RequestBatcher  # noqa: E402
from low_level_feature_extraction_b200.synth import design_image  # noqa: E402


def png_of(img):
    ok, buf = cv2.imencode(".png", img)
    assert ok
    return buf.tobytes()


def check(client, img):
    f = {"file": ("d.png", png_of(img), "image/png")}
    shapes = client.post("/extract-shapes", files=f)
    assert shapes.status_code == 200 and shapes.json() == ShapeAnalyzer.analyze_shapes(img)
    assert shapes.json() == refpath.analyze_shapes_from_mask(refpath.shape_mask(img), img.shape)   # the reference's own calls
    shadows = client.post("/extract-shadows", files=f)
    assert shadows.status_code == 200 and shadows.json() == {"shadow_level": ShadowAnalyzer.analyze_shadow_level(img)}
    assert shadows.json()["shadow_level"] == refpath.shadow_level(img)
    colors = client.post("/extract-colors", files=f)
    assert colors.status_code == 200
    body = colors.json()
    assert body["metadata"]["success"] is True and len(body["accent"]) == 3
    for hx in [body["primary"], body["background"]] + body["accent"]:
        assert len(hx) == 7 and hx[0] == "#"


def test_endpoints_without_batcher():
    with TestClient(create_app(None)) as c:
        check(c, design_image(270, 480, 3))


def test_endpoints_with_batcher_and_concurrent_uploads():
    imgs = [design_image(270, 480, s) for s in range(6)]
    with RequestBatcher(device=0, max_batch=8, max_wait_ms=50.0) as rb, TestClient(create_app(rb)) as c:
        check(c, imgs[0])
        before = rb.batches
        out = [None] * len(imgs)
        go = threading.Barrier(len(imgs))

        def client(i):
            go.wait()
            out[i] = c.post("/extract-shapes", files={"file": ("d.png", png_of(imgs[i]), "image/png")}).json()

        th = [threading.Thread(target=client, args=(i,)) for i in range(len(imgs))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert rb.batches - before < len(imgs)                     # launches were shared
        for i, img in enumerate(imgs):
            assert out[i] == ShapeAnalyzer.analyze_shapes(img)


def test_large_upload_is_resized_like_the_reference():
    img = design_image(1200, 2400, 1)                               # > 2000 px: utils.py:120-127 (INTER_AREA to 2000 x 1000)
    with TestClient(create_app(None)) as c:
        r = c.post("/extract-shapes", files={"file": ("d.png", png_of(img), "image/png")})
        assert r.status_code == 200 and r.json()["metadata"] == {"image_width": 2000, "image_height": 1000}
        assert r.json() == ShapeAnalyzer.analyze_shapes(refpath.auto_resize(img))
