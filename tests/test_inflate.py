"""csrc/h_inflate.cu (host-only entry point llfe_inflate_zlib) against zlib itself: every block type and strategy,
capped output, and damaged streams (same accept / reject decision as zlib, never a crash).  CPU only."""
import ctypes as C
import zlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib():
    import low_level_feature_extraction_b200 as pkg
    return pkg.load_library()


def inflate(lib, data: bytes, cap: int):
    out = C.create_string_buffer(max(cap, 1))
    got = C.c_size_t(0)
    rc = lib.llfe_inflate_zlib(data, len(data), out, cap, C.byref(got))
    return rc, out.raw[:got.value]


def payloads():
    rng = np.random.default_rng(0)
    from low_level_feature_extraction_b200.synth import design_image

    img = design_image(120, 200, 1)
    flat = np.zeros(70000, np.uint8)
    flat[::997] = 9
    yield b""
    yield b"a"
    yield bytes(range(256)) * 3
    yield flat.tobytes()                                            # long matches, distance 1
    yield rng.integers(0, 256, 100000, dtype=np.uint8).tobytes()    # incompressible: stored blocks / long codes
    yield rng.integers(0, 4, 50000, dtype=np.uint8).tobytes()       # short codes
    yield (np.cumsum(rng.integers(-1, 2, 80000)) & 255).astype(np.uint8).tobytes()
    yield img.tobytes()
    yield np.diff(img.astype(np.int16), axis=1, prepend=0).astype(np.uint8).tobytes()   # Sub-filtered rows
    yield (b"abcdefg" * 5000) + rng.integers(0, 256, 300, dtype=np.uint8).tobytes() + b"xyz" * 9000


def test_every_strategy_and_level_equals_zlib(lib):
    n = 0
    for data in payloads():
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
                for wbits in (15, 9):
                    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 9 if n % 2 else 1, strategy)
                    z = c.compress(data[:len(data) // 2]) + c.flush(zlib.Z_FULL_FLUSH) + c.compress(data[len(data) // 2:]) + c.flush()
                    rc, out = inflate(lib, z, len(data))
                    assert rc == 0 and out == data, (len(data), level, strategy, wbits)
                    n += 1
    assert n == 400


def test_capped_output_and_exact_fit(lib):
    data = bytes(np.random.default_rng(1).integers(0, 7, 40000, dtype=np.uint8)) + b"q" * 3000
    z = zlib.compress(data, 6)
    for cap in (0, 1, 100, 39999, 40000, 41500, len(data) - 1, len(data), len(data) + 50):
        rc, out = inflate(lib, z, cap)
        assert rc == 0 and out == data[:cap]
    # trailing bytes after the stream are ignored, like zlib's unused_data
    rc, out = inflate(lib, z + b"tail", len(data))
    assert rc == 0 and out == data


def zlib_accepts(z: bytes, n: int) -> bool:
    try:
        d = zlib.decompressobj()
        out = d.decompress(z, n)
        if len(out) < n:
            return False
        return True
    except zlib.error:
        return False


def test_damaged_streams_are_rejected_like_zlib(lib):
    rng = np.random.default_rng(2)
    data = bytes((np.cumsum(rng.integers(-2, 3, 30000)) & 255).astype(np.uint8))
    z = zlib.compress(data, 6)
    # header, truncation, check value
    assert inflate(lib, b"\x79" + z[1:], len(data))[0] != 0
    assert inflate(lib, z[:1], len(data))[0] != 0
    bad = bytearray(z)
    bad[-1] ^= 1
    assert inflate(lib, bytes(bad), len(data))[0] != 0
    for cut in (2, 3, 10, len(z) // 2, len(z) - 5, len(z) - 1):
        rc, out = inflate(lib, z[:cut], len(data))
        assert rc != 0 or len(out) < len(data), cut
    # random corruption: whenever this inflate accepts and fills the output, zlib must produce the same bytes
    agree = 0
    for trial in range(3000):
        b = bytearray(z)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(2, len(b)))] = int(rng.integers(0, 256))
        rc, out = inflate(lib, bytes(b), len(data))
        ok = rc == 0 and len(out) == len(data)
        ref_ok = zlib_accepts(bytes(b), len(data))
        if ok:
            # zlib.decompress with max_length does not verify what follows either: compare contents
            d = zlib.decompressobj()
            try:
                ref = d.decompress(bytes(b), len(data))
            except zlib.error:
                ref = None
            assert ref == out, trial
        else:
            assert not ref_ok or rc != 0, trial
        agree += ok == ref_ok
    assert agree >= 2990    # (a corrupted trailer is seen by this inflate, not by a capped zlib call)


# ---- several decoders on one stream (llfe_inflate_zlib_mt: speculative block starts) ---------------------------------------
def inflate_mt(lib, data: bytes, cap: int, threads: int):
    out = C.create_string_buffer(max(cap, 1))
    got = C.c_size_t(0)
    rc = lib.llfe_inflate_zlib_mt(data, len(data), out, cap, C.byref(got), threads)
    return rc, out.raw[:got.value]


def big_payloads():
    """megabytes of output, so that every decoder gets more than 128 KB of compressed data"""
    rng = np.random.default_rng(7)
    from low_level_feature_extraction_b200.synth import design_image

    img = design_image(540, 960, 2)
    yield "design rows", img.tobytes()
    yield "sub-filtered rows", np.diff(img.astype(np.int16), axis=1, prepend=0).astype(np.uint8).tobytes()
    yield "noise", rng.integers(0, 256, 1_500_000, dtype=np.uint8).tobytes()           # incompressible: stored blocks
    yield "four symbols", rng.integers(0, 4, 6_000_000, dtype=np.uint8).tobytes()
    walk = (np.cumsum(rng.integers(-1, 2, 3_000_000)) & 255).astype(np.uint8)
    yield "random walk", walk.tobytes()
    # long matches whose sources lie far back, across the places where the workers start (window symbols are copied on)
    block = rng.integers(0, 256, 30000, dtype=np.uint8).tobytes()
    yield "repeats at distance 30000", b"".join(block[:int(k)] + bytes([i & 255]) for i, k in enumerate(rng.integers(20000, 30000, 120)))
    yield "mixed", img.tobytes()[:700000] + rng.integers(0, 256, 400000, dtype=np.uint8).tobytes() + walk.tobytes()[:900000] + b"\0" * 500000


@pytest.mark.parametrize("threads", [2, 3, 4, 8])
def test_mt_equals_zlib(lib, threads):
    for name, data in big_payloads():
        for level, strategy in ((1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FILTERED), (3, zlib.Z_RLE),
                                (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_FIXED), (0, zlib.Z_DEFAULT_STRATEGY)):
            c = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
            z = c.compress(data) + c.flush()
            rc, out = inflate_mt(lib, z, len(data), threads)
            assert rc == 0 and out == data, (name, level, strategy, threads)


def test_mt_full_flushes_small_blocks_and_caps(lib):
    rng = np.random.default_rng(8)
    data = (np.cumsum(rng.integers(-2, 3, 2_500_000)) & 255).astype(np.uint8).tobytes()
    # many tiny blocks, sync / full flushes (empty stored blocks) in the stream
    c = zlib.compressobj(6, zlib.DEFLATED, 15, 1)
    z = b""
    for i in range(0, len(data), 50000):
        z += c.compress(data[i:i + 50000]) + c.flush(zlib.Z_SYNC_FLUSH if (i // 50000) % 3 else zlib.Z_FULL_FLUSH)
    z += c.flush()
    assert len(z) > 600000
    for threads in (2, 4, 7):
        rc, out = inflate_mt(lib, z, len(data), threads)
        assert rc == 0 and out == data
        # output capped anywhere: what fits, no error (libpng: too much image data), exactly as the one-decoder call
        for cap in (0, 1, 1000, len(data) // 4, len(data) // 2 + 17, len(data) - 1, len(data) + 99):
            rc, out = inflate_mt(lib, z, cap, threads)
            rc1, out1 = inflate(lib, z, cap)
            assert (rc, out) == (rc1, out1) and out == data[:cap], (threads, cap)


def test_mt_damaged_streams_agree_with_the_single_decoder(lib):
    """whatever several decoders return -- result code and bytes -- is what one decoder returns: truncations, flipped
    bits anywhere (before, at and behind the places the workers start from), a damaged check value"""
    rng = np.random.default_rng(9)
    from low_level_feature_extraction_b200.synth import design_image

    data = design_image(400, 700, 4).tobytes() + rng.integers(0, 8, 900000, dtype=np.uint8).tobytes()
    z = zlib.compress(data, 6)
    assert len(z) > 500000
    cases = [z[:k] for k in (len(z) - 1, len(z) - 4, len(z) - 5, len(z) // 2, len(z) // 4 * 3 + 1, 300000)]
    bad = bytearray(z)
    bad[-1] ^= 1
    cases.append(bytes(bad))
    for _ in range(120):
        b = bytearray(z)
        for _ in range(int(rng.integers(1, 3))):
            b[int(rng.integers(2, len(b)))] ^= 1 << int(rng.integers(0, 8))
        cases.append(bytes(b))
    accepted = 0
    for k, zz in enumerate(cases):
        rc1, out1 = inflate(lib, zz, len(data))
        for threads in (2, 4):
            rc, out = inflate_mt(lib, zz, len(data), threads)
            assert rc == rc1, (k, threads)
            if rc == 0:
                assert out == out1, (k, threads)
        accepted += rc1 == 0
    assert accepted < len(cases)


def test_mt_concurrent_calls(lib):
    """several callers at once (ctypes releases the GIL): the workers of one call never touch another call's state"""
    from concurrent.futures import ThreadPoolExecutor

    rng = np.random.default_rng(10)
    datas = [(np.cumsum(rng.integers(-2, 3, 1_200_000 + 100_000 * k)) & 255).astype(np.uint8).tobytes() for k in range(4)]
    zs = [zlib.compress(d, 6) for d in datas]

    def job(i):
        k = i % 4
        rc, out = inflate_mt(lib, zs[k], len(datas[k]), 2 + i % 3)
        return rc == 0 and out == datas[k]

    with ThreadPoolExecutor(max_workers=4) as ex:
        assert all(ex.map(job, range(24)))


def test_mt_false_block_starts(lib):
    """deflate streams stored inside a deflate stream: the workers find block headers that parse -- and blocks that
    decode, up to a final block -- in what is only payload; nobody ever arrives at such a start, so the result is that
    of the one-decoder call"""
    rng = np.random.default_rng(11)
    inner = b"".join(zlib.compress((np.cumsum(rng.integers(-2, 3, 400_000)) & 255).astype(np.uint8).tobytes(), 6) for _ in range(6))
    assert len(inner) > 1_000_000
    for outer_level in (0, 1, 6):
        z = zlib.compress(inner, outer_level)
        for threads in (2, 4, 8):
            rc, out = inflate_mt(lib, z, len(inner), threads)
            assert rc == 0 and out == inner, (outer_level, threads)
    # the same with real blocks in front and behind (the plain decoder hands over before and after the stored part)
    walk = (np.cumsum(rng.integers(-2, 3, 2_000_000)) & 255).astype(np.uint8).tobytes()
    c = zlib.compressobj(6)
    z = c.compress(walk) + c.flush(zlib.Z_FULL_FLUSH)
    c0 = zlib.compressobj(0)
    mid = c0.compress(inner) + c0.flush(zlib.Z_FULL_FLUSH)
    # (splice the raw deflate data: a zlib header, compressed blocks, stored blocks, compressed blocks, the final block)
    tail = zlib.compressobj(6, zlib.DEFLATED, -15)
    raw = z[2:] + mid[2:] + tail.compress(walk[::-1]) + tail.flush()
    data = walk + inner + walk[::-1]
    full = z[:2] + raw + zlib.adler32(data).to_bytes(4, "big")
    assert zlib.decompress(full) == data
    for threads in (2, 3, 4, 8):
        rc, out = inflate_mt(lib, full, len(data), threads)
        assert rc == 0 and out == data, threads
