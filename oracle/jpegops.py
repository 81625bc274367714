"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of `cv2.imdecode(buf, cv2.IMREAD_COLOR)` for
baseline (sequential, Huffman, 8-bit) JPEG input -- the decode step of the reference's `validate_and_preprocess_image`
(/root/reference/app/services/analyze/utils.py:108-109) and `ImageProcessor.load_cv2_image` (image_processor.py:62-66).

The arithmetic lives in libjpeg-turbo (OpenCV's bundled build: 3.1.2), which is not under /root/reference; restated here from
the published algorithm of its default decompression path and pinned against the installed cv2 binary by
tests/test_oracle_jpeg.py:

  * entropy decoding: ITU T.81 Annex F (Huffman, DC prediction, restart intervals);
  * dequantisation + inverse DCT: jidctint.c `jpeg_idct_islow` (the default JDCT_ISLOW; 13-bit constants, two passes,
    DESCALE with rounding, range limit around 128);
  * chroma up-sampling: jdsample.c "fancy" triangle filters (h2v1: 3/4 + 1/4 with alternating rounding; h2v2: the same
    both ways, 9/16 3/16 3/16 1/16), edge rows / columns replicated;
  * colour conversion: jdcolor.c YCbCr -> RGB with 16-bit fixed-point tables; OpenCV asks for BGR order.
"""
from __future__ import annotations

import struct

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                   28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                   54, 47, 55, 62, 63])


class Huff:
    def __init__(self, counts, symbols):
        self.lookup = {}
        code = 0
        k = 0
        for length in range(1, 17):
            for _ in range(counts[length - 1]):
                self.lookup[(length, code)] = symbols[k]
                code += 1
                k += 1
            code <<= 1


class Bits:
    def __init__(self, data: bytes):
        self.d = data
        self.p = 0
        self.buf = 0
        self.n = 0

    def bit(self) -> int:
        if self.n == 0:
            b = self.d[self.p] if self.p < len(self.d) else 0
            self.p += 1
            if b == 0xFF:
                nxt = self.d[self.p] if self.p < len(self.d) else 0
                if nxt == 0:
                    self.p += 1          # stuffed zero
            self.buf = b
            self.n = 8
        self.n -= 1
        return (self.buf >> self.n) & 1

    def bits(self, k: int) -> int:
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v

    def symbol(self, h: Huff) -> int:
        code = 0
        for length in range(1, 17):
            code = (code << 1) | self.bit()
            s = h.lookup.get((length, code))
            if s is not None:
                return s
        raise ValueError("bad Huffman code")


def extend(v: int, t: int) -> int:
    return v if t == 0 or v >= (1 << (t - 1)) else v - (1 << t) + 1


def parse(buf: bytes):
    """-> dict(width, height, comps=[(id, h, v, tq)], qt={id: (64,) natural order}, dc / ac tables, scan data, ri)"""
    assert buf[:2] == b"\xff\xd8"
    pos = 2
    qt, dc, ac = {}, {}, {}
    frame = None
    ri = 0
    while True:
        assert buf[pos] == 0xFF
        m = buf[pos + 1]
        pos += 2
        if m == 0xD8 or (0xD0 <= m <= 0xD7) or m == 0x01:
            continue
        if m == 0xD9:
            raise ValueError("no scan")
        (n,) = struct.unpack(">H", buf[pos:pos + 2])
        seg = buf[pos + 2:pos + n]
        pos += n
        if m == 0xDB:
            q = 0
            while q < len(seg):
                pq, tq = seg[q] >> 4, seg[q] & 15
                q += 1
                if pq:
                    vals = struct.unpack(">64H", seg[q:q + 128])
                    q += 128
                else:
                    vals = list(seg[q:q + 64])
                    q += 64
                t = np.zeros(64, np.int64)
                t[ZIGZAG] = np.array(vals, np.int64)
                qt[tq] = t
        elif m == 0xC4:
            q = 0
            while q < len(seg):
                tc, th = seg[q] >> 4, seg[q] & 15
                counts = list(seg[q + 1:q + 17])
                nsym = sum(counts)
                syms = list(seg[q + 17:q + 17 + nsym])
                q += 17 + nsym
                (ac if tc else dc)[th] = Huff(counts, syms)
        elif m in (0xC0, 0xC1):
            p, h, w, nc = struct.unpack(">BHHB", seg[:6])
            assert p == 8
            comps = [(seg[6 + 3 * i], seg[7 + 3 * i] >> 4, seg[7 + 3 * i] & 15, seg[8 + 3 * i]) for i in range(nc)]
            frame = (w, h, comps)
        elif m == 0xC2:
            raise ValueError("progressive")
        elif m == 0xDD:
            (ri,) = struct.unpack(">H", seg[:2])
        elif m == 0xDA:
            ns = seg[0]
            sel = {seg[1 + 2 * i]: (seg[2 + 2 * i] >> 4, seg[2 + 2 * i] & 15) for i in range(ns)}
            end = buf.rfind(b"\xff\xd9")
            return {"width": frame[0], "height": frame[1], "comps": frame[2], "qt": qt, "dc": dc, "ac": ac, "sel": sel,
                    "ri": ri, "scan": buf[pos:end if end > pos else len(buf)]}


def entropy_decode(j):
    """-> per component an array (blocks_y, blocks_x, 64) of quantised coefficients in natural order"""
    comps = j["comps"]
    hmax = max(c[1] for c in comps)
    vmax = max(c[2] for c in comps)
    mcux = -(-j["width"] // (8 * hmax))
    mcuy = -(-j["height"] // (8 * vmax))
    if len(comps) == 1:          # a single-component scan is not interleaved: one block per MCU
        hmax = vmax = 1
        comps = [(comps[0][0], 1, 1, comps[0][3])]
        mcux, mcuy = -(-j["width"] // 8), -(-j["height"] // 8)
    coef = [np.zeros((mcuy * c[2], mcux * c[1], 64), np.int64) for c in comps]
    # restart markers split the scan; each interval starts byte-aligned with the predictors reset
    data = j["scan"]
    segs = []
    if j["ri"]:
        cur = bytearray()
        i = 0
        while i < len(data):
            if data[i] == 0xFF and i + 1 < len(data) and 0xD0 <= data[i + 1] <= 0xD7:
                segs.append(bytes(cur))
                cur = bytearray()
                i += 2
                continue
            cur.append(data[i])
            i += 1
        segs.append(bytes(cur))
    else:
        segs = [data]
    mcu = 0
    for seg in segs:
        br = Bits(seg)
        pred = [0] * len(comps)
        count = j["ri"] if j["ri"] else mcux * mcuy
        for _ in range(count):
            if mcu >= mcux * mcuy:
                break
            my, mx = divmod(mcu, mcux)
            for ci, c in enumerate(comps):
                tdc, tac = j["sel"][c[0]]
                for by in range(c[2]):
                    for bx in range(c[1]):
                        blk = coef[ci][my * c[2] + by, mx * c[1] + bx]
                        t = br.symbol(j["dc"][tdc])
                        pred[ci] += extend(br.bits(t), t) if t else 0
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = br.symbol(j["ac"][tac])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = extend(br.bits(s), s)
                            k += 1
            mcu += 1
    return coef, (hmax, vmax, mcux, mcuy), comps


C = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137,
         f1_961=16069, f2_053=16819, f2_562=20995, f3_072=25172)


def _idct_1d(x, shift):
    """jidctint.c: one pass over the last axis of x (..., 8) int64; DESCALE by `shift`."""
    z2, z3 = x[..., 2], x[..., 6]
    z1 = (z2 + z3) * C["f0_541"]
    tmp2 = z1 + z3 * (-C["f1_847"])
    tmp3 = z1 + z2 * C["f0_765"]
    z2, z3 = x[..., 0], x[..., 4]
    tmp0 = (z2 + z3) << 13
    tmp1 = (z2 - z3) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = x[..., 7], x[..., 5], x[..., 3], x[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * C["f1_175"]
    tmp0 = tmp0 * C["f0_298"]
    tmp1 = tmp1 * C["f2_053"]
    tmp2 = tmp2 * C["f3_072"]
    tmp3 = tmp3 * C["f1_501"]
    z1 = z1 * (-C["f0_899"])
    z2 = z2 * (-C["f2_562"])
    z3 = z3 * (-C["f1_961"]) + z5
    z4 = z4 * (-C["f0_390"]) + z5
    tmp0 = tmp0 + z1 + z3
    tmp1 = tmp1 + z2 + z4
    tmp2 = tmp2 + z2 + z3
    tmp3 = tmp3 + z1 + z4
    r = 1 << (shift - 1)
    out = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2,
                    tmp10 - tmp3], -1)
    return (out + r) >> shift


def idct_islow(coef: np.ndarray, q: np.ndarray) -> np.ndarray:
    """(..., 64) quantised coefficients + (64,) quantisation table -> (..., 8, 8) samples 0..255"""
    x = (coef * q).reshape(coef.shape[:-1] + (8, 8))
    ws = _idct_1d(np.swapaxes(x, -1, -2), 13 - 2)            # pass 1: columns
    ws = np.swapaxes(ws, -1, -2)
    y = _idct_1d(ws, 13 + 2 + 3)                             # pass 2: rows
    return np.clip(y + 128, 0, 255)


def plane(coef, q):
    s = idct_islow(coef, q)                                  # (by, bx, 8, 8)
    by, bx = s.shape[:2]
    return s.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)


def h2v1_fancy(p: np.ndarray) -> np.ndarray:
    p = p.astype(np.int64)
    left = np.concatenate([p[:, :1], p[:, :-1]], 1)
    right = np.concatenate([p[:, 1:], p[:, -1:]], 1)
    out = np.empty((p.shape[0], 2 * p.shape[1]), np.int64)
    out[:, 0::2] = (3 * p + left + 1) >> 2
    out[:, 1::2] = (3 * p + right + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out


def h2v2_fancy(p: np.ndarray) -> np.ndarray:
    p = p.astype(np.int64)
    up = np.concatenate([p[:1], p[:-1]], 0)
    dn = np.concatenate([p[1:], p[-1:]], 0)
    out = np.empty((2 * p.shape[0], 2 * p.shape[1]), np.int64)
    for v, other in ((0, up), (1, dn)):
        cs = 3 * p + other                                   # column sums of the two nearest rows
        last = np.concatenate([cs[:, :1], cs[:, :-1]], 1)
        nxt = np.concatenate([cs[:, 1:], cs[:, -1:]], 1)
        a = (3 * cs + last + 8) >> 4
        b = (3 * cs + nxt + 7) >> 4
        a[:, 0] = (4 * cs[:, 0] + 8) >> 4
        b[:, -1] = (4 * cs[:, -1] + 7) >> 4
        out[v::2, 0::2] = a
        out[v::2, 1::2] = b
    return out


def ycc_to_bgr(y, cb, cr):
    def fix(v):
        return int(v * 65536 + 0.5)

    x = np.arange(256, dtype=np.int64) - 128
    cr_r = (fix(1.40200) * x + 32768) >> 16
    cb_b = (fix(1.77200) * x + 32768) >> 16
    cr_g = -fix(0.71414) * x
    cb_g = -fix(0.34414) * x + 32768
    r = np.clip(y + cr_r[cr], 0, 255)
    g = np.clip(y + ((cb_g[cb] + cr_g[cr]) >> 16), 0, 255)
    b = np.clip(y + cb_b[cb], 0, 255)
    return np.stack([b, g, r], -1).astype(np.uint8)


def imdecode_color(buf: bytes) -> np.ndarray:
    j = parse(buf)
    coef, (hmax, vmax, mcux, mcuy), comps = entropy_decode(j)
    w, h = j["width"], j["height"]
    planes = []
    for ci, c in enumerate(comps):
        p = plane(coef[ci], j["qt"][c[3]])
        # the component's real size: ceil(image size * sampling / max sampling); what lies beyond is block padding
        cw = -(-w * c[1] // hmax)
        ch = -(-h * c[2] // vmax)
        p = p[:ch, :cw]
        if c[1] == hmax and c[2] == vmax:
            up = p
        elif c[1] * 2 == hmax and c[2] == vmax:
            up = h2v1_fancy(p)
        elif c[1] * 2 == hmax and c[2] * 2 == vmax:
            up = h2v2_fancy(p)
        else:
            raise ValueError("sampling not restated")
        planes.append(up[:h, :w].astype(np.int64))
    if len(planes) == 1:
        g = planes[0].astype(np.uint8)
        return np.stack([g, g, g], -1)
    return ycc_to_bgr(planes[0], planes[1], planes[2])


# ---- progressive JPEG (SOF2): T.81 Annex G -- several scans fill the same coefficient arrays -----------------------------

def _segments(buf: bytes):
    """yield (marker, segment payload, position after the segment) for the marker segments of the file"""
    pos = 2
    while pos + 4 <= len(buf):
        assert buf[pos] == 0xFF
        while buf[pos] == 0xFF:
            pos += 1
        m = buf[pos]
        pos += 1
        if m == 0xD9:
            return
        if m in (0x01,) or 0xD0 <= m <= 0xD7:
            continue
        (n,) = struct.unpack(">H", buf[pos:pos + 2])
        yield m, buf[pos + 2:pos + n], pos + n
        pos += n
        if m == 0xDA:       # skip the entropy-coded data up to the next marker that is not RSTn / stuffing
            while pos + 1 < len(buf) and not (buf[pos] == 0xFF and buf[pos + 1] != 0 and not 0xD0 <= buf[pos + 1] <= 0xD7):
                pos += 1


def decode_progressive(buf: bytes):
    """-> (coef per component as in entropy_decode, (hmax, vmax, mcux, mcuy), comps, qt, width, height)"""
    qt, dc, ac = {}, {}, {}
    frame = None
    ri = 0
    coef = None
    for m, seg, after in _segments(buf):
        if m == 0xDB:
            q = 0
            while q < len(seg):
                pq, tq = seg[q] >> 4, seg[q] & 15
                q += 1
                vals = struct.unpack(">64H", seg[q:q + 128]) if pq else list(seg[q:q + 64])
                q += 128 if pq else 64
                t = np.zeros(64, np.int64)
                t[ZIGZAG] = np.array(vals, np.int64)
                qt[tq] = t
        elif m == 0xC4:
            q = 0
            while q < len(seg):
                tc, th = seg[q] >> 4, seg[q] & 15
                counts = list(seg[q + 1:q + 17])
                nsym = sum(counts)
                (ac if tc else dc)[th] = Huff(counts, list(seg[q + 17:q + 17 + nsym]))
                q += 17 + nsym
        elif m == 0xC2:
            p, h, w, nc = struct.unpack(">BHHB", seg[:6])
            comps = [(seg[6 + 3 * i], seg[7 + 3 * i] >> 4, seg[7 + 3 * i] & 15, seg[8 + 3 * i]) for i in range(nc)]
            if nc == 1:
                comps = [(comps[0][0], 1, 1, comps[0][3])]
            hmax = max(c[1] for c in comps)
            vmax = max(c[2] for c in comps)
            mcux = -(-w // (8 * hmax))
            mcuy = -(-h // (8 * vmax))
            frame = (w, h, comps, hmax, vmax, mcux, mcuy)
            coef = [np.zeros((mcuy * c[2], mcux * c[1], 64), np.int64) for c in comps]
        elif m == 0xDD:
            (ri,) = struct.unpack(">H", seg[:2])
        elif m == 0xDA:
            w, h, comps, hmax, vmax, mcux, mcuy = frame
            ns = seg[0]
            scan = []
            for i in range(ns):
                cid = seg[1 + 2 * i]
                ci = [c[0] for c in comps].index(cid)
                scan.append((ci, seg[2 + 2 * i] >> 4, seg[2 + 2 * i] & 15))
            ss, se, ahal = seg[1 + 2 * ns], seg[2 + 2 * ns], seg[3 + 2 * ns]
            ah, al = ahal >> 4, ahal & 15
            # entropy data of this scan: up to the next non-RST marker
            end = after
            while end + 1 < len(buf) and not (buf[end] == 0xFF and buf[end + 1] != 0 and not 0xD0 <= buf[end + 1] <= 0xD7):
                end += 1
            data = buf[after:end]
            segs = []
            cur = bytearray()
            i = 0
            while i < len(data):
                if data[i] == 0xFF and i + 1 < len(data) and 0xD0 <= data[i + 1] <= 0xD7:
                    segs.append(bytes(cur))
                    cur = bytearray()
                    i += 2
                    continue
                cur.append(data[i])
                i += 1
            segs.append(bytes(cur))
            # the units of the scan: MCUs (interleaved) or the blocks of the one component (its real size, not MCU-padded)
            if ns > 1:
                units = [(my, mx) for my in range(mcuy) for mx in range(mcux)]
            else:
                c = comps[scan[0][0]]
                bw = -(-(-(-w * c[1] // hmax)) // 8)
                bh = -(-(-(-h * c[2] // vmax)) // 8)
                units = [(by, bx) for by in range(bh) for bx in range(bw)]
            per = ri if ri else len(units)
            ui = 0
            for sdata in segs:
                br = Bits(sdata)
                pred = [0] * len(comps)
                eobrun = 0
                for _ in range(per):
                    if ui >= len(units):
                        break
                    uy, ux = units[ui]
                    ui += 1
                    blocks = []
                    if ns > 1:
                        for (ci, td, ta) in scan:
                            c = comps[ci]
                            for by in range(c[2]):
                                for bx in range(c[1]):
                                    blocks.append((ci, td, ta, coef[ci][uy * c[2] + by, ux * c[1] + bx]))
                    else:
                        ci, td, ta = scan[0]
                        blocks.append((ci, td, ta, coef[ci][uy, ux]))
                    for ci, td, ta, blk in blocks:
                        if ss == 0:
                            if ah == 0:
                                t = br.symbol(dc[td])
                                pred[ci] += extend(br.bits(t), t) if t else 0
                                blk[0] = pred[ci] * (1 << al)
                            elif br.bit():
                                blk[0] |= 1 << al
                            continue
                        if ah == 0:
                            if eobrun:
                                eobrun -= 1
                                continue
                            k = ss
                            while k <= se:
                                rs = br.symbol(ac[ta])
                                r, s = rs >> 4, rs & 15
                                if s == 0:
                                    if r < 15:
                                        eobrun = (1 << r) - 1 + (br.bits(r) if r else 0)
                                        break
                                    k += 16
                                else:
                                    k += r
                                    blk[ZIGZAG[k]] = extend(br.bits(s), s) * (1 << al)
                                    k += 1
                        else:
                            bit = 1 << al

                            def refine(idx):
                                v = blk[idx]
                                if br.bit() and (v & bit) == 0:
                                    blk[idx] = v + bit if v > 0 else v - bit

                            if eobrun:
                                eobrun -= 1
                                for k in range(ss, se + 1):
                                    if blk[ZIGZAG[k]] != 0:
                                        refine(ZIGZAG[k])
                                continue
                            k = ss
                            while k <= se:
                                rs = br.symbol(ac[ta])
                                r, s = rs >> 4, rs & 15
                                val = 0
                                if s == 0:
                                    if r < 15:
                                        eobrun = (1 << r) - 1 + (br.bits(r) if r else 0)
                                        r = 64
                                else:
                                    val = bit if br.bit() else -bit
                                while k <= se:
                                    idx = ZIGZAG[k]
                                    k += 1
                                    if blk[idx] != 0:
                                        refine(idx)
                                    else:
                                        if r == 0:
                                            if s:
                                                blk[idx] = val
                                            break
                                        r -= 1
    w, h, comps, hmax, vmax, mcux, mcuy = frame
    return coef, (hmax, vmax, mcux, mcuy), comps, qt, w, h


def imdecode_color_progressive(buf: bytes) -> np.ndarray:
    coef, (hmax, vmax, mcux, mcuy), comps, qt, w, h = decode_progressive(buf)
    planes = []
    for ci, c in enumerate(comps):
        p = plane(coef[ci], qt[c[3]])
        cw = -(-w * c[1] // hmax)
        ch = -(-h * c[2] // vmax)
        p = p[:ch, :cw]
        if c[1] == hmax and c[2] == vmax:
            up = p
        elif c[1] * 2 == hmax and c[2] == vmax:
            up = h2v1_fancy(p)
        elif c[1] * 2 == hmax and c[2] * 2 == vmax:
            up = h2v2_fancy(p)
        else:
            raise ValueError("sampling not restated")
        planes.append(up[:h, :w].astype(np.int64))
    if len(planes) == 1:
        g = planes[0].astype(np.uint8)
        return np.stack([g, g, g], -1)
    return ycc_to_bgr(planes[0], planes[1], planes[2])
