"""Routers (services/api.py): wiring and the reference's error mapping, without a GPU."""
import numpy as np
import pytest

fastapi = pytest.importorskip("fastapi")
from fastapi.testclient import TestClient  # noqa: E402

from low_level_feature_extraction_b200.services.api import create_app  # noqa: E402
from low_level_feature_extraction_b200.services import ColorFeatures  # noqa: E402


class FakeBatcher:
    n_colors = 5
    shapes = True

    def __init__(self):
        self.seen = []
        self.closed = False

    def analyze(self, image):
        self.seen.append(image.shape)
        return {"colors": ColorFeatures(primary="#112233", background="#FFFFFF", accent=["#000000"] * 3, metadata={"success": True}),
                "shapes": {"shapes": [], "total_shapes": 0, "metadata": {"image_width": image.shape[1], "image_height": image.shape[0]}},
                "shadow_level": "Low"}

    def close(self):
        self.closed = True


def png(h=40, w=60):
    """An upload that decodes without a GPU: PNG bytes go through the device decoder (services/png.py, covered by
    tests/test_gpu_decode.py), every other format through cv2.imdecode -- so the wiring tests here upload BMP."""
    import cv2

    img = np.random.default_rng(0).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ok, buf = cv2.imencode(".bmp", img)
    assert ok
    return buf.tobytes()


def test_undecodable_upload_is_a_400_like_the_reference():
    with TestClient(create_app(None)) as c:
        for path in ("/extract-colors", "/extract-shapes", "/extract-shadows"):
            r = c.post(path, files={"file": ("x.png", b"this is not an image", "image/png")})
            assert r.status_code == 400, (path, r.text)          # utils.py:111-115 / 147-152
        r = c.post("/extract-colors?preprocessing=nope", files={"file": ("x.png", png(), "image/png")})
        assert r.status_code == 400
        r = c.post("/extract-colors?n_colors=0", files={"file": ("x.png", png(), "image/png")})
        assert r.status_code == 422
        assert c.get("/").json()["endpoints"] == ["/extract-colors", "/extract-shapes", "/extract-shadows"]


def test_requests_go_through_the_batcher_and_keep_the_json_shapes():
    fb = FakeBatcher()
    with TestClient(create_app(fb)) as c:
        r = c.post("/extract-colors", files={"file": ("x.png", png(), "image/png")})
        assert r.status_code == 200
        body = r.json()
        assert body["primary"] == "#112233" and body["background"] == "#FFFFFF" and len(body["accent"]) == 3
        assert body["metadata"] == {"success": True}
        r = c.post("/extract-shapes", files={"file": ("x.png", png(30, 50), "image/png")})
        assert r.status_code == 200 and r.json()["metadata"] == {"image_width": 50, "image_height": 30}
        r = c.post("/extract-shadows", files={"file": ("x.png", png(), "image/png")})
        assert r.status_code == 200 and r.json() == {"shadow_level": "Low"}
    assert fb.seen == [(40, 60, 3), (30, 50, 3), (40, 60, 3)] and fb.closed
