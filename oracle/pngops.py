"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of `cv2.imdecode(buf, cv2.IMREAD_COLOR)` for
non-interlaced PNG input -- the decode step of the reference's `validate_and_preprocess_image`
(/root/reference/app/services/analyze/utils.py:108-109) and `ImageProcessor.load_cv2_image` / `auto_process_image`
(image_processor.py:62-66, :208-211).

The arithmetic lives in two third-party libraries that are not under /root/reference: libpng 1.6.53 (scanline
reconstruction, PNG specification section 9: filters None / Sub / Up / Average / Paeth) and OpenCV's grfmt_png.cpp
(the libpng transforms it requests for IMREAD_COLOR: palette -> RGB, gray 1/2/4 -> 8 bits, strip 16 -> 8 = the high
byte, strip alpha, gray -> colour, RGB -> BGR).  Pinned against the installed cv2 binary by tests/test_oracle_png.py
(every colour type / bit depth / filter type, files written by Pillow, by cv2.imencode and by the writer below).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}


def chunks(buf: bytes):
    pos = 8
    while pos + 12 <= len(buf):
        (n,) = struct.unpack_from(">I", buf, pos)
        yield buf[pos + 4:pos + 8], buf[pos + 8:pos + 8 + n]
        pos += 12 + n


ADAM7 = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]   # x0, y0, dx, dy


def adam7_passes(w: int, h: int):
    """-> [(x0, y0, dx, dy, pass width, pass height)] of the non-empty passes (PNG specification 8.2)"""
    out = []
    for x0, y0, dx, dy in ADAM7:
        pw = (w - x0 + dx - 1) // dx if w > x0 else 0
        ph = (h - y0 + dy - 1) // dy if h > y0 else 0
        if pw and ph:
            out.append((x0, y0, dx, dy, pw, ph))
    return out


def parse(buf: bytes):
    """-> (w, h, depth, color_type, palette bytes, scanline stream)"""
    assert buf[:8] == b"\x89PNG\r\n\x1a\n"
    ihdr, pal, idat = None, b"", []
    for t, d in chunks(buf):
        if t == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", d)
        elif t == b"PLTE":
            pal = d
        elif t == b"IDAT":
            idat.append(d)
    w, h, depth, color, _, _, interlace = ihdr
    parse.interlace = interlace
    return w, h, depth, color, pal, zlib.decompress(b"".join(idat))


def paeth(a: int, b: int, c: int) -> int:
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def unfilter(stream: bytes, h: int, rowbytes: int, bpp: int) -> np.ndarray:
    """PNG specification 9.2: Recon(x) = Filt(x) + predictor, bytes `bpp` apart.  -> (h, rowbytes) uint8"""
    raw = np.frombuffer(stream, np.uint8)[:h * (rowbytes + 1)].reshape(h, rowbytes + 1)
    out = np.zeros((h, rowbytes), np.uint8)
    prev = np.zeros(rowbytes, np.int64)
    for y in range(h):
        ft = int(raw[y, 0])
        f = raw[y, 1:].astype(np.int64)
        if ft == 0:
            cur = f
        elif ft == 1:
            cur = f.copy()
            for k in range(bpp):
                cur[k::bpp] = np.cumsum(f[k::bpp]) & 255
        elif ft == 2:
            cur = (f + prev) & 255
        elif ft == 3:
            cur = np.zeros(rowbytes, np.int64)
            for i in range(rowbytes):
                a = cur[i - bpp] if i >= bpp else 0
                cur[i] = (f[i] + ((a + prev[i]) >> 1)) & 255
        elif ft == 4:
            cur = np.zeros(rowbytes, np.int64)
            for i in range(rowbytes):
                a = int(cur[i - bpp]) if i >= bpp else 0
                c = int(prev[i - bpp]) if i >= bpp else 0
                cur[i] = (f[i] + paeth(a, int(prev[i]), c)) & 255
        else:
            raise ValueError("bad adaptive filter value")
        out[y] = cur
        prev = cur
    return out


def samples(rows: np.ndarray, w: int, channels: int, depth: int) -> np.ndarray:
    """(h, rowbytes) bytes -> (h, w, channels) integer samples, 16-bit ones reduced to their high byte (strip_16)."""
    h = rows.shape[0]
    if depth == 8:
        return rows[:, :w * channels].reshape(h, w, channels).astype(np.int64)
    if depth == 16:
        return rows[:, :2 * w * channels:2].reshape(h, w, channels).astype(np.int64)
    bits = np.unpackbits(rows, axis=1)[:, :w * depth].reshape(h, w, depth).astype(np.int64)
    v = np.zeros((h, w), np.int64)
    for k in range(depth):
        v = (v << 1) | bits[:, :, k]
    return v[:, :, None]


def to_bgr(rows: np.ndarray, w: int, color: int, depth: int, palette: bytes) -> np.ndarray:
    s = samples(rows, w, CHANNELS[color], depth)
    if color == 0:
        g = s[:, :, 0] * {1: 255, 2: 85, 4: 17}.get(depth, 1)
        rgb = np.stack([g, g, g], 2)
    elif color == 2:
        rgb = s
    elif color == 3:
        pal = np.zeros((256, 3), np.int64)
        p = np.frombuffer(palette, np.uint8).reshape(-1, 3)
        pal[:len(p)] = p
        rgb = pal[s[:, :, 0]]
    elif color == 4:
        rgb = np.repeat(s[:, :, :1], 3, 2)
    else:
        rgb = s[:, :, :3]
    return np.ascontiguousarray(rgb[:, :, ::-1]).astype(np.uint8)


def imdecode_color(buf: bytes) -> np.ndarray:
    w, h, depth, color, pal, stream = parse(buf)
    bpp = max(1, CHANNELS[color] * depth // 8)
    if not parse.interlace:
        rowbytes = (w * CHANNELS[color] * depth + 7) // 8
        return to_bgr(unfilter(stream, h, rowbytes, bpp), w, color, depth, pal)
    # Adam7: seven reduced images, each filtered on its own, one after the other in the stream
    out = np.zeros((h, w, 3), np.uint8)
    off = 0
    for x0, y0, dx, dy, pw, ph in adam7_passes(w, h):
        rb = (pw * CHANNELS[color] * depth + 7) // 8
        rows = unfilter(stream[off:off + ph * (rb + 1)], ph, rb, bpp)
        off += ph * (rb + 1)
        out[y0::dy, x0::dx] = to_bgr(rows, pw, color, depth, pal)
    return out


# ---- a PNG writer with a chosen filter per row (test input: Pillow / OpenCV pick their own) ---------------------------------

def _chunk(t: bytes, d: bytes) -> bytes:
    return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))


def filter_rows(rows: np.ndarray, bpp: int, filters) -> bytes:
    """(h, rowbytes) uint8 + one filter type per row -> scanline stream"""
    h, rb = rows.shape
    out = bytearray()
    prev = np.zeros(rb, np.int64)
    for y in range(h):
        cur = rows[y].astype(np.int64)
        a = np.concatenate([np.zeros(bpp, np.int64), cur[:-bpp]]) if rb > bpp else np.zeros(rb, np.int64)
        c = np.concatenate([np.zeros(bpp, np.int64), prev[:-bpp]]) if rb > bpp else np.zeros(rb, np.int64)
        ft = int(filters[y])
        if ft == 0:
            pred = np.zeros(rb, np.int64)
        elif ft == 1:
            pred = a
        elif ft == 2:
            pred = prev
        elif ft == 3:
            pred = (a + prev) >> 1
        else:
            p = a + prev - c
            pa, pb, pc = np.abs(p - a), np.abs(p - prev), np.abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
        out.append(ft)
        out += ((cur - pred) & 255).astype(np.uint8).tobytes()
        prev = cur
    return bytes(out)


def pack_samples(samples: np.ndarray, depth: int) -> np.ndarray:
    """(rows, n) integer samples of `depth` bits -> (rows, ceil(n * depth / 8)) packed bytes (MSB first)"""
    r, n = samples.shape
    if depth == 8:
        return samples.astype(np.uint8)
    if depth == 16:
        o = np.zeros((r, 2 * n), np.uint8)
        o[:, 0::2] = samples >> 8
        o[:, 1::2] = samples & 255
        return o
    bits = np.zeros((r, n * depth), np.uint8)
    for k in range(depth):
        bits[:, k::depth] = (samples >> (depth - 1 - k)) & 1
    pad = (-bits.shape[1]) % 8
    if pad:
        bits = np.concatenate([bits, np.zeros((r, pad), np.uint8)], 1)
    return np.packbits(bits, axis=1)


def write_png_interlaced(samples: np.ndarray, color: int, depth: int, rng, palette: bytes = b"", level: int = 6) -> bytes:
    """samples: (h, w, channels) integers of `depth` bits -> an Adam7 PNG with random filter types per pass row"""
    h, w, ch = samples.shape
    bpp = max(1, ch * depth // 8)
    stream = b""
    for x0, y0, dx, dy, pw, ph in adam7_passes(w, h):
        sub = samples[y0::dy, x0::dx].reshape(ph, pw * ch)
        stream += filter_rows(pack_samples(sub, depth), bpp, rng.integers(0, 5, ph))
    out = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color, 0, 0, 1))
    if palette:
        out += _chunk(b"PLTE", palette)
    return out + _chunk(b"IDAT", zlib.compress(stream, level)) + _chunk(b"IEND", b"")


def write_png(rows: np.ndarray, w: int, color: int, depth: int, filters, palette: bytes = b"", extra=(), level: int = 6,
              idat_split: int = 0) -> bytes:
    """rows: (h, rowbytes) packed scanline bytes.  extra: (type, data) chunks placed before IDAT."""
    h = rows.shape[0]
    bpp = max(1, CHANNELS[color] * depth // 8)
    z = zlib.compress(filter_rows(rows, bpp, filters), level)
    out = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color, 0, 0, 0))
    if palette:
        out += _chunk(b"PLTE", palette)
    for t, d in extra:
        out += _chunk(t, d)
    if idat_split:
        for i in range(0, len(z), idat_split):
            out += _chunk(b"IDAT", z[i:i + idat_split])
    else:
        out += _chunk(b"IDAT", z)
    return out + _chunk(b"IEND", b"")
