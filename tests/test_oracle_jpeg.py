"""The baseline-JPEG restatement (oracle/jpegops.py) against the installed cv2 binary, and the library's host-side entropy
decoder (llfe_jpeg_info / llfe_jpeg_coefficients, host-only entry points) against the oracle's -- CPU only."""
import ctypes as C

import cv2
import numpy as np
import pytest

from oracle import jpegops


def cv2_decode(buf):
    return cv2.imdecode(np.frombuffer(buf, np.uint8), cv2.IMREAD_COLOR)


def encode(img, *params):
    ok, enc = cv2.imencode(".jpg", img, list(params))
    assert ok
    return enc.tobytes()


def cases():
    from low_level_feature_extraction_b200.synth import design_image

    rng = np.random.default_rng(0)
    a = design_image(64, 96, 1)
    b = design_image(61, 83, 2)
    noise = rng.integers(0, 256, (40, 56, 3), dtype=np.uint8)
    yield "420 q95", encode(a)
    yield "420 q50", encode(a, cv2.IMWRITE_JPEG_QUALITY, 50)
    yield "420 odd size", encode(b, cv2.IMWRITE_JPEG_QUALITY, 75)
    yield "gray", encode(cv2.cvtColor(b, cv2.COLOR_BGR2GRAY))
    yield "444", encode(b, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444)
    yield "422", encode(b, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422)
    yield "noise q90", encode(noise, cv2.IMWRITE_JPEG_QUALITY, 90)
    yield "noise q100", encode(noise, cv2.IMWRITE_JPEG_QUALITY, 100)
    yield "restart 3", encode(b, cv2.IMWRITE_JPEG_RST_INTERVAL, 3)
    yield "optimised tables", encode(b, cv2.IMWRITE_JPEG_OPTIMIZE, 1)
    yield "1 x 1", encode(noise[:1, :1])
    yield "17 x 9", encode(noise[:9, :17], cv2.IMWRITE_JPEG_QUALITY, 30)
    P = cv2.IMWRITE_JPEG_PROGRESSIVE
    yield "progressive 420", encode(a, P, 1)
    yield "progressive odd q50", encode(b, P, 1, cv2.IMWRITE_JPEG_QUALITY, 50)
    yield "progressive gray", encode(cv2.cvtColor(b, cv2.COLOR_BGR2GRAY), P, 1)
    yield "progressive 444", encode(b, P, 1, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444)
    yield "progressive 422 restart", encode(b, P, 1, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
                                            cv2.IMWRITE_JPEG_RST_INTERVAL, 3)
    yield "progressive noise q100", encode(noise, P, 1, cv2.IMWRITE_JPEG_QUALITY, 100)
    yield "progressive optimised", encode(b, P, 1, cv2.IMWRITE_JPEG_OPTIMIZE, 1)


CASES = list(cases())


@pytest.mark.parametrize("name,buf", CASES, ids=[c[0] for c in CASES])
def test_oracle_equals_cv2(name, buf):
    dec = jpegops.imdecode_color_progressive if name.startswith("progressive") else jpegops.imdecode_color
    assert np.array_equal(dec(buf), cv2_decode(buf))


@pytest.fixture(scope="module")
def lib():
    import low_level_feature_extraction_b200 as pkg
    return pkg.load_library()


@pytest.mark.parametrize("name,buf", CASES, ids=[c[0] for c in CASES])
def test_native_entropy_decoder_equals_the_oracle(lib, name, buf):
    info = (C.c_int32 * 2)()
    assert lib.llfe_jpeg_info(buf, len(buf), info) == 0
    ref = cv2_decode(buf)
    assert (info[0], info[1]) == (ref.shape[1], ref.shape[0])
    n = C.c_size_t(0)
    assert lib.llfe_jpeg_coefficients(buf, len(buf), None, 0, C.byref(n)) == 0
    out = np.zeros(n.value, np.int16)
    assert lib.llfe_jpeg_coefficients(buf, len(buf), out.ctypes.data, n.value, C.byref(n)) == 0
    if name.startswith("progressive"):
        coef = jpegops.decode_progressive(buf)[0]
    else:
        coef, _, _ = jpegops.entropy_decode(jpegops.parse(buf))
    want = np.concatenate([c.reshape(-1) for c in coef]).astype(np.int16)
    assert np.array_equal(out, want)


def test_files_outside_the_subset_are_refused(lib):
    from low_level_feature_extraction_b200.synth import design_image

    img = design_image(40, 56, 3)
    info = (C.c_int32 * 2)()
    base = encode(img)
    prog = encode(img, cv2.IMWRITE_JPEG_PROGRESSIVE, 1)
    # a progression that stops early (libjpeg would smooth the blocks): cut the file after its third scan and close it
    sos = [i for i in range(len(prog) - 1) if prog[i] == 0xFF and prog[i + 1] == 0xDA]
    assert len(sos) > 4 and lib.llfe_jpeg_info(prog, len(prog), info) == 0
    early = prog[:sos[3]] + b"\xff\xd9"
    n0 = C.c_size_t(0)
    lib.llfe_jpeg_coefficients(prog, len(prog), None, 0, C.byref(n0))
    scratch = np.zeros(n0.value, np.int16)
    assert lib.llfe_jpeg_coefficients(early, len(early), scratch.ctypes.data, n0.value, C.byref(n0)) == -4   # LLFE_E_UNSUPPORTED
    exif = base[:2] + b"\xff\xe1\x00\x10Exif\x00\x00" + b"\0" * 8 + base[2:]      # an Exif segment: OpenCV may rotate
    assert lib.llfe_jpeg_info(exif, len(exif), info) == -4
    assert lib.llfe_jpeg_info(b"\x89PNG\r\n\x1a\n" + b"0" * 32, 40, info) == -1  # not a JPEG
    assert lib.llfe_jpeg_info(base[:100], 100, info) == -1                       # header cut off
    # damaged entropy data never crashes: an error or some coefficients
    rng = np.random.default_rng(1)
    n = C.c_size_t(0)
    lib.llfe_jpeg_coefficients(base, len(base), None, 0, C.byref(n))
    out = np.zeros(n.value, np.int16)
    for src in (base, prog):
        for trial in range(400):
            b = bytearray(src)
            lo = 2 if trial % 4 == 0 else len(b) // 3          # every fourth trial may hit the headers too
            for _ in range(3):
                b[int(rng.integers(lo, len(b)))] = int(rng.integers(0, 256))
            cnt = C.c_size_t(0)
            rc = lib.llfe_jpeg_coefficients(bytes(b), len(b), None, 0, C.byref(cnt))
            assert rc in (0, -1, -4)
            if rc == 0 and cnt.value == n.value:                # same geometry: decode into the buffer
                rc = lib.llfe_jpeg_coefficients(bytes(b), len(b), out.ctypes.data, n.value, C.byref(cnt))
                assert rc in (0, -1, -4)
