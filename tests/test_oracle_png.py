"""The PNG restatement (oracle/pngops.py) against the installed cv2 binary, and the product's chunk parser
(services/png.py: which files the device path takes) -- CPU only, no compute calls."""
import io
import struct
import zlib

import cv2
import numpy as np
import pytest
from PIL import Image

from oracle import pngops

CASES = [  # (colour type, bit depth)
    (0, 1), (0, 2), (0, 4), (0, 8), (0, 16), (2, 8), (2, 16), (3, 1), (3, 2), (3, 4), (3, 8), (4, 8), (4, 16), (6, 8), (6, 16),
]


def cv2_decode(buf):
    return cv2.imdecode(np.frombuffer(buf, np.uint8), cv2.IMREAD_COLOR)


def make_case(color, depth, h, w, seed, filters=None, **kw):
    rng = np.random.default_rng(seed)
    rb = (w * pngops.CHANNELS[color] * depth + 7) // 8
    rows = rng.integers(0, 256, (h, rb), dtype=np.uint8)
    if (w * pngops.CHANNELS[color] * depth) % 8:   # padding bits of the last byte are zero in well-formed files
        rows[:, -1] &= (0xFF << (8 - (w * pngops.CHANNELS[color] * depth) % 8)) & 0xFF
    # smooth rows make the Average / Paeth predictors take every branch
    rows[h // 2:] = (np.cumsum(rng.integers(-2, 3, (h - h // 2, rb)), axis=1) + 128).astype(np.uint8)
    palette = b""
    if color == 3:
        n = min(1 << depth, int(rng.integers(2, 257)))
        palette = rng.integers(0, 256, n * 3, dtype=np.uint8).tobytes()
        if depth == 8:
            rows = (rows % n).astype(np.uint8)   # libpng rejects nothing here, but keep the indices meaningful
    if filters is None:
        filters = rng.integers(0, 5, h)
    return pngops.write_png(rows, w, color, depth, filters, palette, **kw)


@pytest.mark.parametrize("color,depth", CASES)
def test_oracle_equals_cv2_on_every_colour_type_and_filter(color, depth):
    for seed, (h, w) in enumerate([(13, 29), (7, 1), (1, 9), (40, 67)]):
        buf = make_case(color, depth, h, w, 100 * color + depth + seed)
        ref = cv2_decode(buf)
        assert ref is not None
        assert np.array_equal(pngops.imdecode_color(buf), ref), (color, depth, h, w)


@pytest.mark.parametrize("ft", [0, 1, 2, 3, 4])
def test_oracle_equals_cv2_single_filter(ft):
    buf = make_case(2, 8, 21, 33, ft, filters=[ft] * 21)
    assert np.array_equal(pngops.imdecode_color(buf), cv2_decode(buf))
    buf = make_case(6, 8, 21, 33, ft, filters=[ft] * 21, idat_split=100)
    assert np.array_equal(pngops.imdecode_color(buf), cv2_decode(buf))


def test_oracle_equals_cv2_on_pillow_and_opencv_files():
    from low_level_feature_extraction_b200.synth import design_image

    img = design_image(54, 96, 3)
    ok, enc = cv2.imencode(".png", img)
    assert np.array_equal(pngops.imdecode_color(enc.tobytes()), img)
    for mode in ("RGB", "RGBA", "L", "LA", "P"):
        b = io.BytesIO()
        Image.fromarray(img[:, :, ::-1]).convert(mode).save(b, "PNG", optimize=True)
        assert np.array_equal(pngops.imdecode_color(b.getvalue()), cv2_decode(b.getvalue())), mode


def test_product_parser_accepts_plain_pngs_and_defers_the_rest():
    from low_level_feature_extraction_b200.services import png

    buf = make_case(2, 8, 9, 11, 1, extra=[(b"gAMA", struct.pack(">I", 45455)), (b"tEXt", b"k\0v")], idat_split=17)
    info = png.parse(buf)
    assert (info.width, info.height, info.bit_depth, info.color_type) == (11, 9, 8, 2)
    w, h, depth, color, pal, stream = pngops.parse(buf)
    assert png.inflate(info) == stream and info.rowbytes == 33
    # deferred to cv2.imdecode: not a PNG, unknown interlace method, APNG / unknown chunk, damaged CRC, truncated file, short stream
    assert png.parse(b"\xff\xd8\xff\xe0" + b"0" * 64) is None
    inter = bytearray(buf)
    inter[28] = 2
    inter[29:33] = struct.pack(">I", zlib.crc32(bytes(inter[12:29])))
    assert png.parse(bytes(inter)) is None
    assert png.parse(make_case(2, 8, 9, 11, 1, extra=[(b"acTL", b"\0" * 8)])) is None
    bad = bytearray(buf)
    bad[40] ^= 1
    assert png.parse(bytes(bad)) is None
    assert png.parse(buf[:-20]) is None
    short = png.parse(pngops.write_png(np.zeros((4, 33), np.uint8), 11, 2, 8, [0] * 4))
    short.height = 5
    assert png.inflate(short) is None
    pal = make_case(3, 4, 6, 7, 2)
    assert len(png.parse(pal).palette) % 3 == 0 and png.parse(pal).rowbytes == 4


ADAM7_SIZES = [(1, 1), (3, 5), (8, 8), (9, 17), (40, 67), (2, 1), (1, 7), (5, 2)]


def make_adam7_case(color, depth, h, w, seed):
    """an Adam7-interlaced file with random samples and a random filter type per pass row"""
    rng = np.random.default_rng(seed)
    hi, pal = 1 << depth, b""
    if color == 3:
        hi = min(hi, int(rng.integers(2, 257)))
        pal = rng.integers(0, 256, hi * 3, dtype=np.uint8).tobytes()
    s = rng.integers(0, hi, (h, w, pngops.CHANNELS[color]))
    s[h // 2:] = (np.cumsum(rng.integers(-2, 3, s[h // 2:].shape), axis=1) + hi // 2) % hi   # smooth rows: every predictor branch
    return pngops.write_png_interlaced(s, color, depth, rng, pal)


@pytest.mark.parametrize("color,depth", CASES)
def test_oracle_equals_cv2_on_adam7_files(color, depth):
    """PNG specification 8.2: seven reduced images, each filtered on its own (passes narrower than the image are skipped)"""
    from low_level_feature_extraction_b200.services import png

    for seed, (h, w) in enumerate(ADAM7_SIZES):
        buf = make_adam7_case(color, depth, h, w, 1000 * color + 10 * depth + seed)
        ref = cv2_decode(buf)
        assert ref is not None
        assert np.array_equal(pngops.imdecode_color(buf), ref), (color, depth, h, w)
        info = png.parse(buf)      # the product's parser takes the file and sizes the stream like the inflate does
        assert info is not None and info.interlace == 1 and info.stream_bytes == len(pngops.parse(buf)[5])


def test_parser_and_inflate_never_raise_on_damaged_files():
    """Every damaged file is either refused by the chunk walk (CRC, structure) or by the inflate -- the caller then
    leaves it to cv2.imdecode; nothing raises and nothing reads out of bounds."""
    from low_level_feature_extraction_b200.services import png

    rng = np.random.default_rng(4)
    base = [make_case(2, 8, 31, 45, 1), make_case(6, 8, 12, 19, 2, idat_split=50), make_case(3, 4, 20, 33, 3), make_case(0, 16, 9, 14, 4)]
    refused = accepted = 0
    for trial in range(600):
        b = bytearray(base[trial % len(base)])
        mode = trial % 3
        if mode == 0:      # flip bytes anywhere (almost always caught by a chunk CRC)
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] ^= int(rng.integers(1, 256))
        elif mode == 1:    # truncate
            b = b[:int(rng.integers(0, len(b)))]
        else:              # damage the zlib stream and repair the chunk CRC, so that the inflate sees it
            pos = bytes(b).find(b"IDAT")
            (n,) = struct.unpack(">I", bytes(b[pos - 4:pos]))
            if n > 8:
                b[pos + 4 + int(rng.integers(2, n))] ^= int(rng.integers(1, 256))
                b[pos + 4 + n:pos + 8 + n] = struct.pack(">I", zlib.crc32(bytes(b[pos:pos + 4 + n])))
        info = png.parse(bytes(b))
        if info is None:
            refused += 1
            continue
        stream = png.inflate(info)
        if stream is None:
            refused += 1
        else:
            accepted += 1
            assert len(stream) == info.stream_bytes
    assert refused > 400 and accepted + refused == 600
