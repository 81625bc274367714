"""The Pillow restatement (oracle/pilops.py: thumbnail size, ImagingReduce, ImagingResample with LANCZOS) against the
installed Pillow binary -- the arithmetic of the reference's ImageProcessor.auto_process_image
(app/services/analyze/image_processor.py:221-224).  CPU only."""
import numpy as np
import pytest
from PIL import Image

from oracle import pilops


def pil_thumb(a, mw, mh):
    im = Image.fromarray(a)
    im.thumbnail((mw, mh), Image.Resampling.LANCZOS)
    return np.array(im)


@pytest.mark.parametrize("h,w,mw,mh", [
    (108, 192, 96, 54), (300, 500, 64, 64), (1000, 37, 50, 50), (501, 733, 40, 40), (90, 160, 192, 108), (77, 1200, 100, 100),
    (1300, 9, 30, 30), (640, 480, 31, 57), (450, 450, 33, 20), (216, 384, 192, 108), (217, 383, 192, 108)])
def test_thumbnail_equals_pillow(h, w, mw, mh):
    a = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(pilops.thumbnail_lanczos(a, mw, mh), pil_thumb(a, mw, mh))


def test_resample_with_boxes_up_scaling_and_gray():
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (90, 120, 3), dtype=np.uint8)
    for box in [(0, 0, 120, 90), (1.5, 2.25, 100.75, 80.5), (10, 10, 50, 50), (0.3, 0, 120, 90)]:
        ref = np.array(Image.fromarray(a).resize((40, 30), Image.Resampling.LANCZOS, box=box, reducing_gap=None))
        assert np.array_equal(pilops.resample_lanczos(a, 40, 30, box), ref), box
    ref = np.array(Image.fromarray(a).resize((150, 100), Image.Resampling.LANCZOS))
    assert np.array_equal(pilops.resize_lanczos(a, (150, 100)), ref)
    g = rng.integers(0, 256, (70, 90), dtype=np.uint8)
    assert np.array_equal(pilops.thumbnail_lanczos(g, 31, 29), pil_thumb(g, 31, 29))
    # saturating content: overshoot of the negative lobes must clip like clip8
    s = np.zeros((64, 64, 3), np.uint8)
    s[:, 32:] = 255
    s[20:30] = 255
    assert np.array_equal(pilops.thumbnail_lanczos(s, 20, 20), pil_thumb(s, 20, 20))


@pytest.mark.parametrize("fx,fy", [(2, 2), (3, 3), (4, 4), (5, 5), (2, 3), (7, 5), (1, 4), (4, 1), (9, 7), (16, 16)])
def test_reduce_equals_pillow(fx, fy):
    rng = np.random.default_rng(fx * 31 + fy)
    a = rng.integers(0, 256, (61, 83, 3), dtype=np.uint8)
    a[:20] = 255
    assert np.array_equal(pilops.reduce_box_mean(a, fx, fy), np.array(Image.fromarray(a).reduce((fx, fy))))
    box = (3, 5, 80, 58)
    assert np.array_equal(pilops.reduce_box_mean(a, fx, fy, box), np.array(Image.fromarray(a).reduce((fx, fy), box=box)))


def test_drop_in_thumbnail_host_logic_without_gpu(monkeypatch):
    """services/image_processor.py mirrors PIL/Image.py's size / reduce-factor / box rules on the host; the two device
    primitives are stood in for by the oracle here, so the logic around them is checked against Pillow on CPU."""
    from low_level_feature_extraction_b200.services import image_processor as ip

    def fake_call(src, fx, fy, reduce_box, box, dw, dh):
        img = src
        if fx > 1 or fy > 1:
            img = pilops.reduce_box_mean(img, fx, fy, tuple(int(v) for v in reduce_box))
        return pilops.resample_lanczos(img, dw, dh, tuple(float(np.float32(v)) for v in box))

    monkeypatch.setattr(ip, "_pil_resize_call", fake_call)
    rng = np.random.default_rng(11)
    for h, w, mw, mh in [(108, 192, 96, 54), (300, 500, 64, 64), (501, 733, 40, 40), (90, 160, 192, 108), (77, 1200, 100, 100),
                         (1300, 9, 30, 30), (640, 480, 31, 57), (900, 1400, 100, 75), (1210, 11, 40, 40)]:
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(ip.pil_thumbnail_lanczos(a, mw, mh), pil_thumb(a, mw, mh)), (h, w, mw, mh)
    a = rng.integers(0, 256, (200, 300, 3), dtype=np.uint8)
    for size, box in [((70, 50), None), ((33, 21), (10.5, 20.25, 250.0, 180.5)), ((300, 200), None), ((20, 20), (0, 0, 300, 200))]:
        ref = np.array(Image.fromarray(a).resize(size, Image.Resampling.LANCZOS, box=box, reducing_gap=2.0))
        assert np.array_equal(ip.pil_resize_lanczos(a, size, box, 2.0), ref), (size, box)
    with pytest.raises(TypeError):
        ip.pil_resize_lanczos(a.astype(np.float32), (10, 10))
    with pytest.raises(ValueError):
        ip.pil_resize_lanczos(a, (10, 10), reducing_gap=0.5)
