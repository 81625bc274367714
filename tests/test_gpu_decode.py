"""SURVEY 8(f)3 on the device: PNG reconstruction (`cv2.imdecode(buf, IMREAD_COLOR)` after the host inflate) and Pillow's
LANCZOS thumbnail, through the C ABI, against cv2 / Pillow themselves and the oracle."""
import asyncio
import io

import cv2
import numpy as np
import pytest
import torch
from PIL import Image

from oracle import pilops, pngops
from test_oracle_png import ADAM7_SIZES, CASES, cv2_decode, make_adam7_case, make_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(llfe):
    return llfe.engine(0)


@pytest.mark.parametrize("color,depth", CASES)
def test_png_decode_equals_cv2(color, depth):
    from low_level_feature_extraction_b200.services import png

    for seed, (h, w) in enumerate([(13, 29), (7, 1), (1, 9), (40, 67), (300, 1037)]):
        buf = make_case(color, depth, h, w, 100 * color + depth + seed)
        got = png.decode(buf)
        assert got is not None
        assert np.array_equal(got, cv2_decode(buf)), (color, depth, h, w)


@pytest.mark.parametrize("color,depth", CASES)
def test_adam7_png_decode_equals_cv2(color, depth):
    from low_level_feature_extraction_b200.services import png

    for seed, (h, w) in enumerate(ADAM7_SIZES + [(300, 1037)]):
        buf = make_adam7_case(color, depth, h, w, 1000 * color + 10 * depth + seed)
        got = png.decode(buf)
        assert got is not None
        assert np.array_equal(got, cv2_decode(buf)), (color, depth, h, w)


def test_adam7_png_1080p_in_a_batch_and_damaged():
    from low_level_feature_extraction_b200.services import png
    from low_level_feature_extraction_b200.synth import design_image

    img = design_image(1080, 1920, 5)
    buf = pngops.write_png_interlaced(img[:, :, ::-1].astype(np.int64), 2, 8, np.random.default_rng(0))
    plain = cv2.imencode(".png", img)[1].tobytes()
    out = png.decode_many([plain, buf, buf[:len(buf) // 2]])
    assert np.array_equal(out[0], img) and np.array_equal(out[1], img) and out[2] is None
    # a filter byte of 5 in pass 3: refused (the caller leaves the file to OpenCV)
    w, h, depth, color, pal, stream = pngops.parse(buf)
    s = bytearray(stream)
    off = sum(ph * (3 * pw + 1) for _, _, _, _, pw, ph in pngops.adam7_passes(w, h)[:2])
    s[off] = 5
    import zlib as _z
    bad = buf[:33] + pngops._chunk(b"IDAT", _z.compress(bytes(s), 1)) + pngops._chunk(b"IEND", b"")
    assert png.parse(bad) is not None and png.decode(bad) is None


def test_png_decode_1080p_files_from_pillow_and_opencv():
    from low_level_feature_extraction_b200.services import png
    from low_level_feature_extraction_b200.synth import design_image

    img = design_image(1080, 1920, 11)
    ok, enc = cv2.imencode(".png", img)                       # every row Sub-filtered
    assert np.array_equal(png.decode(enc.tobytes()), img)
    for mode in ("RGB", "RGBA", "L", "P"):                    # Pillow: adaptive filters (Up / Average / Paeth rows)
        b = io.BytesIO()
        Image.fromarray(img[:, :, ::-1]).convert(mode).save(b, "PNG")
        ref = cv2_decode(b.getvalue())
        assert np.array_equal(png.decode(b.getvalue()), ref), mode
        assert np.array_equal(png.imdecode_color(b.getvalue()), ref)
    # tall and narrow / wide and flat: more rows than threads, fewer chunks than a warp
    # ... and rows too long for the shared-memory delay line (the first thread reads the row above from the stream)
    for h, w in [(3000, 5), (3, 9000), (2161, 777), (300, 12500)]:
        rng = np.random.default_rng(h)
        a = (np.cumsum(rng.integers(-3, 4, (h, w, 3)), axis=1) + 100).astype(np.uint8)
        b = io.BytesIO()
        Image.fromarray(a).save(b, "PNG")
        assert np.array_equal(png.decode(b.getvalue()), a[:, :, ::-1]), (h, w)


def test_png_batch_on_device_and_bad_filter_status(eng):
    h, w = 57, 131
    bufs = [make_case(6, 8, h, w, s) for s in range(5)]
    streams = np.stack([np.frombuffer(pngops.parse(b)[5], np.uint8) for b in bufs]).copy()
    streams[3, 5 * (w * 4 + 1)] = 7                           # image 3: invalid filter type in row 5
    out, status = eng.png_reconstruct(torch.from_numpy(streams).cuda(), h, w, 6, 8)
    status = status.cpu().numpy()
    assert list(status != 0) == [False, False, False, True, False]
    for i in (0, 1, 2, 4):
        assert np.array_equal(out[i].cpu().numpy(), cv2_decode(bufs[i])), i
    # colour type 3 with per-image palettes
    bufs = [make_case(3, 8, h, w, 50 + s) for s in range(3)]
    parsed = [pngops.parse(b) for b in bufs]
    pal = np.zeros((3, 256, 3), np.uint8)
    for i, p in enumerate(parsed):
        q = np.frombuffer(p[4], np.uint8).reshape(-1, 3)
        pal[i, :len(q)] = q
    streams = np.stack([np.frombuffer(p[5], np.uint8) for p in parsed]).copy()
    out, status = eng.png_reconstruct(torch.from_numpy(streams).cuda(), h, w, 3, 8, torch.from_numpy(pal).cuda())
    assert not status.any()
    for i in range(3):
        assert np.array_equal(out[i].cpu().numpy(), cv2_decode(bufs[i])), i


def test_decode_many_and_damaged_streams():
    from low_level_feature_extraction_b200.services import png

    bufs = [make_case(c, d, 50 + 7 * i, 81 + i, 300 + i) for i, (c, d) in enumerate(CASES)]
    damaged = bytearray(bufs[2])
    damaged[len(damaged) // 2] ^= 0x55          # inside IDAT: the chunk CRC no longer matches
    bufs.append(bytes(damaged))
    bufs.append(b"not a png")
    got = png.decode_many(bufs, workers=4)
    for b, g in zip(bufs[:len(CASES)], got):
        assert np.array_equal(g, cv2_decode(b))
    assert got[-1] is None and got[-2] is None
    # a valid container around a damaged zlib stream (CRC recomputed): this decoder defers, and OpenCV's answer is returned
    import struct
    import zlib
    rows = np.random.default_rng(3).integers(0, 256, (30, 3 * 40), dtype=np.uint8)
    good = pngops.write_png(rows, 40, 2, 8, [1] * 30)
    z = bytearray(next(d for t, d in pngops.chunks(good) if t == b"IDAT"))
    z[len(z) // 2] ^= 0x10
    bad = good[:33] + struct.pack(">I", len(z)) + b"IDAT" + bytes(z) + struct.pack(">I", zlib.crc32(b"IDAT" + bytes(z))) + good[-12:]
    assert png.parse(bad) is not None
    ref, mine = cv2_decode(bad), png.imdecode_color(bad)
    assert (ref is None and mine is None) or np.array_equal(ref, mine)


def test_bad_filter_byte_is_left_to_opencv():
    from low_level_feature_extraction_b200.services import png

    rows = np.zeros((4, 9), np.uint8)
    buf = pngops.write_png(rows, 3, 2, 8, [0, 9, 0, 0])
    assert png.parse(buf) is not None and png.decode(buf) is None
    ref = cv2_decode(buf)
    got = png.imdecode_color(buf)
    assert (ref is None and got is None) or np.array_equal(ref, got)


def test_validate_and_preprocess_image_on_png_bytes():
    from low_level_feature_extraction_b200.services.utils import HTTPException, validate_and_preprocess_image
    from low_level_feature_extraction_b200.synth import design_image

    img = design_image(300, 2400, 5)
    b = io.BytesIO()
    Image.fromarray(img[:, :, ::-1]).save(b, "PNG")
    buf = b.getvalue()
    got = asyncio.run(validate_and_preprocess_image(buf, "r", "none"))
    assert np.array_equal(got, img)
    got = asyncio.run(validate_and_preprocess_image(buf, "r", "auto"))
    assert np.array_equal(got, cv2.resize(img, (2000, 250), interpolation=cv2.INTER_AREA))
    ok, jpg = cv2.imencode(".jpg", img)
    got = asyncio.run(validate_and_preprocess_image(jpg.tobytes(), "r", "none"))
    assert np.array_equal(got, cv2_decode(jpg.tobytes()))
    with pytest.raises(HTTPException) as e:
        asyncio.run(validate_and_preprocess_image(b"\x89PNG\r\n\x1a\n" + b"garbage" * 10, "r", "auto"))
    assert e.value.status_code == 400


def pil_thumb(a, mw, mh):
    im = Image.fromarray(a)
    im.thumbnail((mw, mh), Image.Resampling.LANCZOS)
    return np.array(im)


@pytest.mark.parametrize("h,w,mw,mh", [
    (108, 192, 96, 54), (300, 500, 64, 64), (1000, 37, 50, 50), (501, 733, 40, 40), (90, 160, 192, 108), (77, 1200, 100, 100),
    (1300, 9, 30, 30), (640, 480, 31, 57), (2160, 3840, 1920, 1080), (2161, 3839, 1920, 1080), (1080, 1920, 1920, 1080),
    (4320, 7680, 1280, 720), (5000, 4000, 999, 999)])
def test_pil_thumbnail_equals_pillow(h, w, mw, mh):
    from low_level_feature_extraction_b200.services.image_processor import pil_thumbnail_lanczos

    a = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    got = pil_thumbnail_lanczos(a, mw, mh)
    assert np.array_equal(got, pil_thumb(a, mw, mh))
    if h * w < 400 * 600:
        assert np.array_equal(got, pilops.thumbnail_lanczos(a, mw, mh))


def test_pil_primitives_on_device(eng):
    rng = np.random.default_rng(9)
    a = rng.integers(0, 256, (2, 90, 120, 3), dtype=np.uint8)
    d = torch.from_numpy(a).cuda()
    for box in [None, (1.5, 2.25, 100.75, 80.5), (10, 10, 50, 50)]:
        got = eng.pil_resample_lanczos(d, 30, 40, box).cpu().numpy()
        for i in range(2):
            ref = np.array(Image.fromarray(a[i]).resize((40, 30), Image.Resampling.LANCZOS, box=box, reducing_gap=None))
            assert np.array_equal(got[i], ref), box
    up = eng.pil_resample_lanczos(d[0], 100, 150).cpu().numpy()
    assert np.array_equal(up, np.array(Image.fromarray(a[0]).resize((150, 100), Image.Resampling.LANCZOS, reducing_gap=None)))
    g = torch.from_numpy(a[0, :, :, 0].copy()).cuda()
    assert np.array_equal(eng.pil_resample_lanczos(g, 29, 31).cpu().numpy(),
                          np.array(Image.fromarray(a[0, :, :, 0]).resize((31, 29), Image.Resampling.LANCZOS, reducing_gap=None)))
    for fx, fy, box in [(2, 2, None), (3, 5, None), (4, 1, (3, 5, 80, 58)), (7, 7, (0, 0, 120, 90))]:
        got = eng.pil_reduce(d, fx, fy, box).cpu().numpy()
        for i in range(2):
            assert np.array_equal(got[i], np.array(Image.fromarray(a[i]).reduce((fx, fy), box=box))), (fx, fy)
    # saturating content (negative lobes overshoot) and the only-one-axis paths
    s = np.zeros((64, 64, 3), np.uint8)
    s[:, 32:] = 255
    s[20:30] = 255
    ds = torch.from_numpy(s).cuda()
    for size in [(20, 20), (64, 20), (20, 64)]:
        ref = np.array(Image.fromarray(s).resize(size, Image.Resampling.LANCZOS, reducing_gap=None))
        assert np.array_equal(eng.pil_resample_lanczos(ds, size[1], size[0]).cpu().numpy(), ref), size


def test_auto_process_image_equals_the_reference_call_sequence():
    from low_level_feature_extraction_b200.services.image_processor import ImageProcessor
    from low_level_feature_extraction_b200.synth import design_image

    img = design_image(1200, 2100, 8)
    ok, enc = cv2.imencode(".png", img)
    buf = enc.tobytes()
    # image_processor.py:208-225, literally
    image = cv2.imdecode(np.frombuffer(buf, np.uint8), cv2.IMREAD_COLOR)
    pil_image = Image.fromarray(cv2.cvtColor(image, cv2.COLOR_BGR2RGB))
    pil_image.thumbnail((1920, 1080), Image.Resampling.LANCZOS)
    ref = cv2.cvtColor(np.array(pil_image), cv2.COLOR_RGB2BGR)
    assert np.array_equal(ImageProcessor.auto_process_image(buf), ref)
    assert np.array_equal(ImageProcessor.load_cv2_image(buf), img)
    with pytest.raises(ValueError):
        ImageProcessor.auto_process_image(b"nonsense")


# ---- baseline JPEG: entropy decoding on the host, IDCT / up-sampling / colour conversion on the device -------------------------
from test_oracle_jpeg import CASES as JPEG_CASES  # noqa: E402
from test_oracle_jpeg import encode as jpeg_encode  # noqa: E402


@pytest.mark.parametrize("name,buf", JPEG_CASES, ids=[c[0] for c in JPEG_CASES])
def test_jpeg_decode_equals_cv2(name, buf):
    from low_level_feature_extraction_b200.services import jpeg

    got = jpeg.decode(buf)
    assert got is not None
    assert np.array_equal(got, cv2_decode(buf))


def test_jpeg_decode_large_files_and_every_sampling():
    from low_level_feature_extraction_b200.services import jpeg, png
    from low_level_feature_extraction_b200.synth import design_image, noise_image

    for h, w, seed in [(1080, 1920, 3), (2160, 3840, 4), (1081, 1919, 5), (7, 2500, 6), (2500, 7, 7)]:
        img = design_image(h, w, seed) if min(h, w) > 16 else noise_image(h, w, seed)
        for params in ([], [cv2.IMWRITE_JPEG_QUALITY, 70], [cv2.IMWRITE_JPEG_PROGRESSIVE, 1, cv2.IMWRITE_JPEG_QUALITY, 85],
                       [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444],
                       [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_RST_INTERVAL, 7]):
            buf = jpeg_encode(img, *params)
            ref = cv2_decode(buf)
            got = jpeg.decode(buf)
            assert got is not None and np.array_equal(got, ref), (h, w, params)
            assert np.array_equal(png.imdecode_color(buf), ref)
    gray = cv2.cvtColor(design_image(333, 517, 8), cv2.COLOR_BGR2GRAY)
    buf = jpeg_encode(gray)
    assert np.array_equal(jpeg.decode(buf), cv2_decode(buf))


def test_jpeg_files_outside_the_subset_go_to_opencv():
    from low_level_feature_extraction_b200.services import jpeg, png
    from low_level_feature_extraction_b200.synth import design_image

    img = design_image(200, 300, 9)
    prog = jpeg_encode(img, cv2.IMWRITE_JPEG_PROGRESSIVE, 1)
    assert np.array_equal(jpeg.decode(prog), cv2_decode(prog))          # progressive files are decoded here too
    sos = [i for i in range(len(prog) - 1) if prog[i] == 0xFF and prog[i + 1] == 0xDA]
    early = prog[:sos[3]] + b"\xff\xd9"                                   # an incomplete progression is OpenCV's call
    assert jpeg.decode(early) is None
    ref, got = cv2_decode(early), png.imdecode_color(early)
    assert (ref is None and got is None) or np.array_equal(ref, got)
    base = jpeg_encode(img)
    cut = base[:len(base) * 2 // 3]                       # truncated: libjpeg pads with gray, this decoder defers
    assert jpeg.decode(cut) is None
    ref, got = cv2_decode(cut), png.imdecode_color(cut)
    assert (ref is None and got is None) or np.array_equal(ref, got)
    rng = np.random.default_rng(5)
    for _ in range(40):                                   # damaged entropy data: whatever OpenCV makes of it
        b = bytearray(base)
        for _ in range(2):
            b[int(rng.integers(len(b) // 2, len(b) - 2))] = int(rng.integers(0, 256))
        mine = jpeg.decode(bytes(b))
        if mine is not None:
            assert np.array_equal(mine, cv2_decode(bytes(b)))
