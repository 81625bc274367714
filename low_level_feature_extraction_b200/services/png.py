"""`cv2.imdecode(buf, cv2.IMREAD_COLOR)` for PNG input with everything after the inflate on the GPU
(reference call sites: app/services/analyze/utils.py:108-109, image_processor.py:62-66, :208-211).

A PNG is a chunk container around ONE zlib stream.  The host walks the chunks (CRC-checked, like libpng does for the
critical ones) and inflates the IDAT stream -- a serial bit-level decode of a single stream, so it stays on a host
core, with the library's own inflate (csrc/h_inflate.cu, 1.1-1.4x zlib 1.2.11) writing straight into pinned memory;
`inflate_many` spreads a batch over threads (the call releases the GIL).  Scanline reconstruction (the five
PNG filters) and the conversion OpenCV asks libpng for (palette / gray expansion, 16 -> 8 bits, alpha dropped,
RGB -> BGR) run in `llfe_png_reconstruct*` (csrc/k_png.cu).

`parse` returns None for anything the device path does not take -- not a PNG, APNG, an unknown chunk, a damaged
file -- and the callers hand those buffers to `cv2.imdecode` exactly as the reference does, so the
decision "is this decodable, and to what" stays OpenCV's for every input outside the plain-PNG case.
"""
from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

from . import _runtime

SIGNATURE = b"\x89PNG\r\n\x1a\n"
_CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}
_DEPTHS = {0: (1, 2, 4, 8, 16), 2: (8, 16), 3: (1, 2, 4, 8), 4: (8, 16), 6: (8, 16)}
# ancillary chunks that do not change the samples libpng hands to OpenCV for IMREAD_COLOR (probed against cv2 4.13:
# gAMA is not applied, tRNS is dropped with the alpha channel)
_HARMLESS = {b"gAMA", b"cHRM", b"sRGB", b"iCCP", b"pHYs", b"tEXt", b"zTXt", b"iTXt", b"tIME", b"bKGD", b"sBIT", b"hIST",
             b"sPLT", b"eXIf", b"tRNS"}
_MAX_PIXELS = 1 << 30   # OpenCV's CV_IO_MAX_IMAGE_PIXELS


@dataclass
class PngInfo:
    width: int
    height: int
    bit_depth: int
    color_type: int
    palette: bytes          # RGB triples (colour type 3), else b""
    idat: bytes             # the concatenated IDAT payloads = one zlib stream
    interlace: int = 0      # IHDR interlace method: 0, or 1 = Adam7

    @property
    def rowbytes(self) -> int:
        return (self.width * _CHANNELS[self.color_type] * self.bit_depth + 7) // 8

    @property
    def stream_bytes(self) -> int:
        if not self.interlace:
            return self.height * (self.rowbytes + 1)
        from .._native import load_library

        return int(load_library().llfe_png_stream_bytes(self.width, self.height, self.color_type, self.bit_depth, 1))


def parse(buf) -> PngInfo | None:
    """Chunk walk of a PNG file; None = not a file the device path takes (see the module docstring)."""
    b = bytes(buf) if not isinstance(buf, (bytes, bytearray, memoryview)) else buf
    mv = memoryview(b)
    n = len(mv)
    if n < 8 + 25 or bytes(mv[:8]) != SIGNATURE:
        return None
    pos = 8
    ihdr = None
    palette = b""
    idat = []
    seen_iend = False
    idat_done = False
    while pos + 12 <= n:
        (length,) = struct.unpack_from(">I", mv, pos)
        ctype = bytes(mv[pos + 4:pos + 8])
        end = pos + 8 + length
        if length > 0x7FFFFFFF or end + 4 > n:
            return None
        data = mv[pos + 8:end]
        (crc,) = struct.unpack_from(">I", mv, end)
        if zlib.crc32(data, zlib.crc32(ctype)) != crc:
            return None
        pos = end + 4
        if ihdr is None:
            if ctype != b"IHDR" or length != 13:
                return None
            ihdr = struct.unpack(">IIBBBBB", data)
            continue
        if ctype == b"IDAT":
            if idat_done:
                return None          # IDAT chunks must be consecutive
            idat.append(data)
        else:
            if idat:
                idat_done = True
            if ctype == b"IEND":
                seen_iend = True
                break
            if ctype == b"PLTE":
                if palette or idat or length == 0 or length % 3 or length > 768:
                    return None
                palette = bytes(data)
            elif ctype not in _HARMLESS:
                return None          # acTL (APNG), unknown or private chunks: OpenCV's call
    if ihdr is None or not seen_iend or not idat:
        return None
    w, h, depth, color, comp, filt, interlace = ihdr
    if color not in _CHANNELS or depth not in _DEPTHS[color] or comp != 0 or filt != 0 or interlace not in (0, 1):
        return None
    if w == 0 or h == 0 or w * h > _MAX_PIXELS or h > 65535:
        return None
    if color == 3 and not palette:
        return None
    return PngInfo(w, h, depth, color, palette, b"".join(idat), interlace)


def inflate(info: PngInfo) -> bytes | None:
    """The scanline stream (height x [filter byte + rowbytes]) via the library's inflate (csrc/h_inflate.cu; the GIL is
    released during the call); None when the zlib stream is damaged or short."""
    import ctypes as C

    from .._native import load_library

    want = info.stream_bytes
    out = C.create_string_buffer(want)
    got = C.c_size_t(0)
    rc = load_library().llfe_inflate_zlib(info.idat, len(info.idat), out, want, C.byref(got))
    if rc != 0 or got.value != want:
        return None
    return out.raw


def inflate_many(infos, workers: int = 8):
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(inflate, infos))


def decode(buf) -> np.ndarray | None:
    """(H, W, 3) uint8 BGR, equal to cv2.imdecode(buf, cv2.IMREAD_COLOR), for the PNGs `parse` accepts; None = hand the
    buffer to cv2.imdecode (not such a PNG, or a stream zlib / libpng would reject)."""
    info = parse(buf)
    if info is None:
        return None
    out = np.empty((info.height, info.width, 3), np.uint8)
    from .._native import LLFE_E_INVALID, LlfeError

    with _runtime.lock():
        try:
            _runtime.context().call("llfe_png_decode_adam7_host" if info.interlace else "llfe_png_decode_host", info.idat,
                                    len(info.idat), info.height, info.width, info.color_type, info.bit_depth,
                                    info.palette or None, len(info.palette) // 3, out)
        except LlfeError as e:
            if e.code == LLFE_E_INVALID:
                return None
            raise
    return out


def decode_many(bufs, workers: int = 8) -> list:
    """`decode` for a batch of files: the inflates run on `workers` host threads (one stream each), the device
    reconstruction follows image by image as the streams arrive.  None entries = hand that buffer to cv2.imdecode."""
    from .._native import LLFE_E_INVALID, LlfeError

    infos = [parse(b) for b in bufs]
    out = [None] * len(bufs)
    for k, info in enumerate(infos):
        if info is not None and info.interlace:      # Adam7: the one-call path
            out[k] = decode(bufs[k])
            infos[k] = None
    with ThreadPoolExecutor(max_workers=workers) as ex:
        futures = [ex.submit(inflate, i) if i is not None else None for i in infos]
        for k, (info, fut) in enumerate(zip(infos, futures)):
            stream = fut.result() if fut is not None else None
            if stream is None:
                continue
            img = np.empty((info.height, info.width, 3), np.uint8)
            with _runtime.lock():
                try:
                    _runtime.context().call("llfe_png_reconstruct_host", stream, info.height, info.width, info.color_type,
                                            info.bit_depth, info.palette or None, len(info.palette) // 3, img)
                except LlfeError as e:
                    if e.code == LLFE_E_INVALID:
                        continue
                    raise
            out[k] = img
    return out


def imdecode_color(buf) -> np.ndarray | None:
    """Drop-in for `cv2.imdecode(np.frombuffer(buf, np.uint8), cv2.IMREAD_COLOR)`: PNG and baseline JPEG take the device
    paths (this module, services/jpeg.py), everything else -- and every file those refuse -- goes to OpenCV."""
    import cv2

    from . import jpeg

    arr = np.frombuffer(buf, np.uint8) if isinstance(buf, (bytes, bytearray, memoryview)) else np.asarray(buf, np.uint8)
    raw = None
    if arr.size >= 8 and arr[:8].tobytes() == SIGNATURE:
        raw = buf if isinstance(buf, (bytes, bytearray)) else arr.tobytes()
        img = decode(raw)
        if img is not None:
            return img
    elif arr.size >= 4 and arr[:2].tobytes() == jpeg.SIGNATURE:
        raw = buf if isinstance(buf, (bytes, bytearray)) else arr.tobytes()
        img = jpeg.decode(raw)
        if img is not None:
            return img
    return cv2.imdecode(arr, cv2.IMREAD_COLOR)
