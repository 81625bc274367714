// Baseline and progressive JPEG -> the BGR image `cv2.imdecode(buf, cv2.IMREAD_COLOR)` returns (reference call sites:
// app/services/analyze/utils.py:108-109, image_processor.py:62-66, :208-211).
//
// Like a PNG's inflate, a JPEG's entropy-coded segment is one serial bit-level decode (every code's position depends on
// all codes before it) and stays on a host core (ITU T.81 Annex F: Huffman codes, DC prediction, restart intervals;
// Annex G for progressive files: spectral selection and successive approximation over several scans);
// it writes the quantised coefficients straight into pinned memory.  Everything after it runs on the device and follows
// libjpeg-turbo's default decompression path bit for bit (restated in oracle/jpegops.py, pinned against cv2):
//
//   k_jpeg_idct     dequantisation + jidctint.c `jpeg_idct_islow` (13-bit constants, two passes, DESCALE with rounding,
//                   range limit around 128): a thread per 8 x 8 block, all 64 values in registers
//   k_jpeg_to_bgr   jdsample.c "fancy" chroma up-sampling (h2v1: 3/4 + 1/4 with alternating rounding, h2v2: the same
//                   both ways; edge rows / columns replicated) + jdcolor.c YCbCr -> RGB in 16-bit fixed point, BGR out
//
// Files outside this subset -- arithmetic-coded, 12-bit, CMYK / Adobe-marked, sequential files with more than one scan,
// progressions that stop before every coefficient is complete (libjpeg smooths those), sampling other than 4:4:4 /
// 4:2:2 / 4:2:0, an Exif orientation to apply -- are refused (LLFE_E_UNSUPPORTED) and the
// caller hands them to cv2.imdecode, as it does with every file this decoder finds damaged.
#include <string.h>

#include <atomic>
#include <thread>

#include "llfe_common.cuh"
#include "llfe_device.cuh"

namespace {

const uint8_t JZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct JHuff {
    bool present = false;
    uint16_t fast[512];      // (length << 8) | symbol for codes of up to 9 bits, 0 = longer code
    int16_t fast_ac[512];    // AC tables: (value << 8) | (run << 4) | total bits when code + magnitude bits fit in 9 bits, else 0
    int32_t maxcode[18];     // largest code of each length (-1 = none), left-aligned compare value per length
    int32_t valptr[17];
    int32_t mincode[17];
    uint8_t vals[256];
};

struct JComp {
    int id, h, v, tq, td, ta;
    int bw, bh;              // blocks per row / column in the coefficient buffer (MCU-padded)
    int cw, ch;              // real size of the component plane in samples
    size_t coef_off;         // in int16 units
    size_t plane_off;        // in bytes (IDCT output planes, bw * 8 wide)
};

struct JInfo {
    int width = 0, height = 0, ncomp = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0, ri = 0;
    JComp comp[3];
    uint16_t qt[4][64];      // natural order
    bool qt_present[4] = {false, false, false, false};
    JHuff dc[4], ac[4];
    const uint8_t* scan = nullptr;
    const uint8_t* end = nullptr;
    size_t coef_count = 0, plane_bytes = 0;
    bool progressive = false;
    size_t first_sos = 0;    // offset of the length field of the first SOS segment
};

bool build_huff(const uint8_t* counts, const uint8_t* vals, int nvals, JHuff* h) {
    h->present = true;
    memset(h->fast, 0, sizeof h->fast);
    memcpy(h->vals, vals, nvals);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
        h->valptr[len] = k;
        h->mincode[len] = code;
        for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
            if (code >= (1 << len)) return false;   // over-subscribed
            if (len <= 9) {
                const int lo = code << (9 - len), n = 1 << (9 - len);
                for (int j = 0; j < n; ++j) h->fast[lo + j] = (uint16_t)((len << 8) | vals[k]);
            }
        }
        h->maxcode[len] = counts[len - 1] ? code - 1 : -1;
        code <<= 1;
    }
    h->maxcode[17] = 0x7fffffff;
    // most AC coefficients of a photograph are small and have short codes: one look-up then yields run, value and length
    for (int i = 0; i < 512; ++i) {
        h->fast_ac[i] = 0;
        const uint16_t f = h->fast[i];
        if (!f) continue;
        const int len = f >> 8, run = (f >> 4) & 15, size = f & 15;
        if (size == 0 || len + size > 9) continue;
        int v = ((i << len) & 511) >> (9 - size);
        if (v < (1 << (size - 1))) v += 1 - (1 << size);
        if (v >= -128 && v <= 127) h->fast_ac[i] = (int16_t)((v * 256) + (run * 16) + (len + size));
    }
    return k == nvals;
}

// JFIF header walk.  Returns LLFE_OK, LLFE_E_UNSUPPORTED (a valid file outside the subset) or LLFE_E_INVALID.
int jpeg_parse(const uint8_t* buf, size_t len, JInfo* J) {
    if (len < 4 || buf[0] != 0xFF || buf[1] != 0xD8) return LLFE_E_INVALID;
    size_t pos = 2;
    bool have_frame = false, jfif = false;
    for (;;) {
        if (pos + 4 > len || buf[pos] != 0xFF) return LLFE_E_INVALID;
        while (pos < len && buf[pos] == 0xFF) ++pos;      // fill bytes
        if (pos >= len) return LLFE_E_INVALID;
        const int m = buf[pos++];
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9) return LLFE_E_INVALID;
        if (pos + 2 > len) return LLFE_E_INVALID;
        const size_t n = ((size_t)buf[pos] << 8) | buf[pos + 1];
        if (n < 2 || pos + n > len) return LLFE_E_INVALID;
        const uint8_t* seg = buf + pos + 2;
        const size_t sl = n - 2;
        pos += n;
        if (m == 0xE0) {
            if (sl >= 5 && !memcmp(seg, "JFIF", 5)) jfif = true;
        } else if (m == 0xE1) {
            if (sl >= 6 && !memcmp(seg, "Exif\0", 6)) return LLFE_E_UNSUPPORTED;   // OpenCV applies the Exif orientation
        } else if (m == 0xEE) {
            if (sl >= 5 && !memcmp(seg, "Adobe", 5)) return LLFE_E_UNSUPPORTED;      // colour transform flag
        } else if (m == 0xDB) {
            size_t q = 0;
            while (q < sl) {
                const int pq = seg[q] >> 4, tq = seg[q] & 15;
                ++q;
                if (tq > 3 || pq > 1 || q + (pq ? 128 : 64) > sl) return LLFE_E_INVALID;
                for (int i = 0; i < 64; ++i) {
                    const int v = pq ? ((seg[q + 2 * i] << 8) | seg[q + 2 * i + 1]) : seg[q + i];
                    J->qt[tq][JZIGZAG[i]] = (uint16_t)v;
                }
                q += pq ? 128 : 64;
                J->qt_present[tq] = true;
            }
        } else if (m == 0xC4) {
            size_t q = 0;
            while (q < sl) {
                if (q + 17 > sl) return LLFE_E_INVALID;
                const int tc = seg[q] >> 4, th = seg[q] & 15;
                if (tc > 1 || th > 3) return LLFE_E_INVALID;
                int nv = 0;
                for (int i = 0; i < 16; ++i) nv += seg[q + 1 + i];
                if (nv > 256 || q + 17 + nv > sl) return LLFE_E_INVALID;
                if (!build_huff(seg + q + 1, seg + q + 17, nv, tc ? &J->ac[th] : &J->dc[th])) return LLFE_E_INVALID;
                q += 17 + nv;
            }
        } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
            if (have_frame || sl < 6) return LLFE_E_INVALID;
            J->progressive = m == 0xC2;
            if (seg[0] != 8) return LLFE_E_UNSUPPORTED;
            J->height = (seg[1] << 8) | seg[2];
            J->width = (seg[3] << 8) | seg[4];
            J->ncomp = seg[5];
            if (J->ncomp != 1 && J->ncomp != 3) return LLFE_E_UNSUPPORTED;
            if (J->width <= 0 || J->height <= 0 || sl < (size_t)(6 + 3 * J->ncomp)) return LLFE_E_INVALID;
            for (int i = 0; i < J->ncomp; ++i) {
                JComp& c = J->comp[i];
                c.id = seg[6 + 3 * i];
                c.h = seg[7 + 3 * i] >> 4;
                c.v = seg[7 + 3 * i] & 15;
                c.tq = seg[8 + 3 * i];
                if (c.tq > 3 || c.h < 1 || c.v < 1) return LLFE_E_INVALID;
            }
            have_frame = true;
        } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return LLFE_E_UNSUPPORTED;   // lossless, hierarchical, arithmetic
        } else if (m == 0xDD) {
            if (sl < 2) return LLFE_E_INVALID;
            J->ri = (seg[0] << 8) | seg[1];
        } else if (m == 0xDA) {
            if (!have_frame || sl < 1) return LLFE_E_INVALID;
            if (J->progressive) {      // the scans are walked by jpeg_progressive_decode
                J->first_sos = pos - n;
                J->end = buf + len;
                break;
            }
            const int ns = seg[0];
            if (ns != J->ncomp || sl < (size_t)(4 + 2 * ns)) return LLFE_E_UNSUPPORTED;   // one interleaved scan only
            for (int i = 0; i < ns; ++i) {
                if (seg[1 + 2 * i] != J->comp[i].id) return LLFE_E_UNSUPPORTED;
                J->comp[i].td = seg[2 + 2 * i] >> 4;
                J->comp[i].ta = seg[2 + 2 * i] & 15;
                if (J->comp[i].td > 3 || J->comp[i].ta > 3) return LLFE_E_INVALID;
                if (!J->dc[J->comp[i].td].present || !J->ac[J->comp[i].ta].present || !J->qt_present[J->comp[i].tq])
                    return LLFE_E_INVALID;
            }
            if (seg[1 + 2 * ns] != 0 || seg[2 + 2 * ns] != 63 || seg[3 + 2 * ns] != 0) return LLFE_E_UNSUPPORTED;
            J->scan = buf + pos;
            J->end = buf + len;
            break;
        }
    }
    // colour space: what libjpeg assumes for a JFIF file / component ids 1, 2, 3
    if (J->ncomp == 3 && !jfif && !(J->comp[0].id == 1 && J->comp[1].id == 2 && J->comp[2].id == 3)) return LLFE_E_UNSUPPORTED;
    if (J->ncomp == 1) {
        J->comp[0].h = J->comp[0].v = 1;        // a single-component scan is not interleaved: one block per MCU
    } else {
        const JComp& y = J->comp[0];
        if (!((y.h == 1 || y.h == 2) && (y.v == 1 || y.v == 2)) || (y.h == 1 && y.v == 2)) return LLFE_E_UNSUPPORTED;
        if (J->comp[1].h != 1 || J->comp[1].v != 1 || J->comp[2].h != 1 || J->comp[2].v != 1) return LLFE_E_UNSUPPORTED;
    }
    J->hmax = J->comp[0].h;
    J->vmax = J->comp[0].v;
    J->mcux = (J->width + 8 * J->hmax - 1) / (8 * J->hmax);
    J->mcuy = (J->height + 8 * J->vmax - 1) / (8 * J->vmax);
    if ((size_t)J->width * J->height > (size_t(1) << 30) || J->height > 65535 * 8) return LLFE_E_UNSUPPORTED;
    size_t co = 0, po = 0;
    for (int i = 0; i < J->ncomp; ++i) {
        JComp& c = J->comp[i];
        c.bw = J->mcux * c.h;
        c.bh = J->mcuy * c.v;
        c.cw = (J->width * c.h + J->hmax - 1) / J->hmax;
        c.ch = (J->height * c.v + J->vmax - 1) / J->vmax;
        c.coef_off = co;
        c.plane_off = po;
        co += (size_t)c.bw * c.bh * 64;
        po += ((size_t)c.bw * 8 * c.bh * 8 + 255) & ~size_t(255);
    }
    J->coef_count = co;
    J->plane_bytes = po;
    for (int i = 0; i < J->ncomp; ++i)
        if (!J->qt_present[J->comp[i].tq]) return LLFE_E_INVALID;
    return LLFE_OK;
}

// MSB-first bit reader over the entropy-coded segment (FF 00 unstuffed; a marker ends the supply of real bits)
struct JBits {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t buf = 0;
    int cnt = 0;
    int marker = 0;      // marker that stopped the reader (0 = none)
    int fake = 0;        // zero bits appended behind the last real one (a marker or the end of the file was reached)
    inline void fill() {
        if (cnt >= 32) return;       // a code (<= 16 bits) and its extra bits (<= 15) are there
        // fast path: four bytes at once when none of them is FF (no stuffing, no marker)
        while (cnt <= 32 && !marker && end - p >= 4) {
            uint32_t v;
            memcpy(&v, p, 4);
            const uint32_t nv = ~v;
            if (((nv - 0x01010101u) & ~nv & 0x80808080u) != 0u) break;      // some byte is FF
            v = __builtin_bswap32(v);
            buf |= (uint64_t)v << (32 - cnt);
            cnt += 32;
            p += 4;
        }
        while (cnt < 32) {
            int b = 0;
            if (marker || p >= end) fake += 8;
            if (!marker && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    const int n = p < end ? *p : 0xD9;
                    if (n == 0) {
                        ++p;
                    } else {
                        marker = n;      // the FF belongs to a marker: feed zeros from here on
                        --p;
                        b = 0;
                        fake += 8;
                    }
                }
            }
            buf |= (uint64_t)b << (56 - cnt);
            cnt += 8;
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf >> (64 - n)); }
    inline void drop(int n) {
        buf <<= n;
        cnt -= n;
    }
};

inline int jdecode(JBits& b, const JHuff& h) {
    const uint16_t f = h.fast[b.peek(9)];
    if (f) {
        b.drop(f >> 8);
        return f & 255;
    }
    int code = (int)b.peek(10);
    for (int len = 10; len <= 16; ++len) {
        if (h.maxcode[len] >= 0 && code <= h.maxcode[len] && code >= h.mincode[len]) {
            b.drop(len);
            return h.vals[h.valptr[len] + code - h.mincode[len]];
        }
        code = (int)b.peek(len + 1);
    }
    return -1;
}

inline int jextend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }

// entropy-coded segment -> quantised coefficients (natural order, int16), blocks [by][bx][64] per component
int jpeg_entropy_decode(const JInfo& J, int16_t* coef, std::atomic<int>* rows_done = nullptr) {
    JBits b;
    b.p = J.scan;
    b.end = J.end;
    int pred[3] = {0, 0, 0};
    const int total = J.mcux * J.mcuy;
    int until_restart = J.ri ? J.ri : total + 1;
    for (int mcu = 0; mcu < total; ++mcu) {
        if (until_restart == 0) {
            // byte-align, expect RSTn
            b.fill();
            if (!(b.marker >= 0xD0 && b.marker <= 0xD7)) return LLFE_E_INVALID;
            b.p += 2;                 // the reader stands on the FF of the marker
            b.buf = 0, b.cnt = 0, b.marker = 0, b.fake = 0;
            pred[0] = pred[1] = pred[2] = 0;
            until_restart = J.ri;
        }
        --until_restart;
        // libjpeg keeps decoding zeros when the data ends early (with a warning); such files are left to it
        if (b.cnt < b.fake) return LLFE_E_INVALID;
        const int my = mcu / J.mcux, mx = mcu - my * J.mcux;
        for (int ci = 0; ci < J.ncomp; ++ci) {
            const JComp& c = J.comp[ci];
            const JHuff& hd = J.dc[c.td];
            const JHuff& ha = J.ac[c.ta];
            for (int by = 0; by < c.v; ++by)
                for (int bx = 0; bx < c.h; ++bx) {
                    int16_t* blk = coef + c.coef_off + ((size_t)(my * c.v + by) * c.bw + (mx * c.h + bx)) * 64;
                    memset(blk, 0, 64 * sizeof(int16_t));     // cleared right before it is written: one pass over the buffer
                    b.fill();
                    const int t = jdecode(b, hd);
                    if (t < 0 || t > 11) return LLFE_E_INVALID;
                    if (t) {
                        pred[ci] += jextend((int)b.peek(t), t);
                        b.drop(t);
                    }
                    blk[0] = (int16_t)pred[ci];
                    for (int k = 1; k < 64;) {
                        b.fill();
                        const int fa = ha.fast_ac[b.peek(9)];
                        if (fa) {
                            k += (fa >> 4) & 15;
                            if (k > 63) return LLFE_E_INVALID;
                            b.drop(fa & 15);
                            blk[JZIGZAG[k++]] = (int16_t)(fa >> 8);
                            continue;
                        }
                        const int rs = jdecode(b, ha);
                        if (rs < 0) return LLFE_E_INVALID;
                        const int r = rs >> 4, s = rs & 15;
                        if (s == 0) {
                            if (r != 15) break;
                            k += 16;
                            continue;
                        }
                        k += r;
                        if (k > 63) return LLFE_E_INVALID;
                        blk[JZIGZAG[k]] = (int16_t)jextend((int)b.peek(s), s);
                        b.drop(s);
                        ++k;
                    }
                }
        }
        if (rows_done && mx == J.mcux - 1) rows_done->store(my + 1, std::memory_order_release);   // this MCU row is final
    }
    if (b.cnt < b.fake) return LLFE_E_INVALID;
    return LLFE_OK;
}

// ---- progressive (SOF2, T.81 Annex G): several scans refine the same coefficient arrays ----------------------------------------
struct JScanComp {
    int ci, td, ta;
};

int jpeg_progressive_decode(JInfo& J, const uint8_t* buf, size_t len, int16_t* coef) {
    memset(coef, 0, J.coef_count * sizeof(int16_t));
    int8_t prec[3][64];               // point transform each coefficient has been sent with so far (-1: not yet)
    memset(prec, -1, sizeof prec);
    size_t pos = J.first_sos;
    for (;;) {
        // ---- scan header
        if (pos + 2 > len) return LLFE_E_INVALID;
        const size_t n = ((size_t)buf[pos] << 8) | buf[pos + 1];
        if (n < 3 || pos + n > len) return LLFE_E_INVALID;
        const uint8_t* seg = buf + pos + 2;
        const int ns = seg[0];
        if (ns < 1 || ns > J.ncomp || n < (size_t)(6 + 2 * ns)) return LLFE_E_INVALID;
        JScanComp sc[3];
        for (int i = 0; i < ns; ++i) {
            int ci = -1;
            for (int c = 0; c < J.ncomp; ++c)
                if (J.comp[c].id == seg[1 + 2 * i]) ci = c;
            if (ci < 0) return LLFE_E_INVALID;
            sc[i] = JScanComp{ci, seg[2 + 2 * i] >> 4, seg[2 + 2 * i] & 15};
            if (sc[i].td > 3 || sc[i].ta > 3) return LLFE_E_INVALID;
        }
        const int ss = seg[1 + 2 * ns], se = seg[2 + 2 * ns], ah = seg[3 + 2 * ns] >> 4, al = seg[3 + 2 * ns] & 15;
        if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss > 0 && ns != 1) || al > 13 || (ah != 0 && ah != al + 1)) return LLFE_E_INVALID;
        for (int i = 0; i < ns; ++i) {
            if (ss == 0 && ah == 0 && !J.dc[sc[i].td].present) return LLFE_E_INVALID;
            if (ss > 0 && !J.ac[sc[i].ta].present) return LLFE_E_INVALID;
            for (int k = ss; k <= se; ++k) {
                int8_t& p = prec[sc[i].ci][k];
                if (ah == 0 ? p != -1 : p != ah) return LLFE_E_UNSUPPORTED;      // an unusual progression: libjpeg's call
                p = (int8_t)al;
            }
        }
        // ---- entropy-coded data of the scan
        JBits b;
        b.p = buf + pos + n;
        b.end = buf + len;
        int pred[3] = {0, 0, 0};
        int eobrun = 0;
        int uw, uh;                   // units of the scan: MCUs, or the blocks of its one component (real size)
        if (ns > 1) {
            uw = J.mcux, uh = J.mcuy;
        } else {
            uw = (J.comp[sc[0].ci].cw + 7) / 8, uh = (J.comp[sc[0].ci].ch + 7) / 8;
        }
        const int total = uw * uh;
        int until_restart = J.ri ? J.ri : total + 1;
        const int bit = 1 << al;
        for (int u = 0; u < total; ++u) {
            if (until_restart == 0) {
                b.fill();
                if (!(b.marker >= 0xD0 && b.marker <= 0xD7)) return LLFE_E_INVALID;
                b.p += 2;
                b.buf = 0, b.cnt = 0, b.marker = 0, b.fake = 0;
                pred[0] = pred[1] = pred[2] = 0;
                eobrun = 0;
                until_restart = J.ri;
            }
            --until_restart;
            if (b.cnt < b.fake) return LLFE_E_INVALID;
            const int uy = u / uw, ux = u - uy * uw;
            for (int i = 0; i < ns; ++i) {
                const JComp& c = J.comp[sc[i].ci];
                const int nv = ns > 1 ? c.v : 1, nh = ns > 1 ? c.h : 1;
                for (int by = 0; by < nv; ++by)
                    for (int bx = 0; bx < nh; ++bx) {
                        int16_t* blk = coef + c.coef_off + ((size_t)(uy * nv + by) * c.bw + (ux * nh + bx)) * 64;
                        if (ss == 0) {
                            b.fill();
                            if (ah == 0) {
                                const int t = jdecode(b, J.dc[sc[i].td]);
                                if (t < 0 || t > 15) return LLFE_E_INVALID;
                                if (t) {
                                    pred[sc[i].ci] += jextend((int)b.peek(t), t);
                                    b.drop(t);
                                }
                                blk[0] = (int16_t)(pred[sc[i].ci] * bit);
                            } else {
                                if (b.peek(1)) blk[0] |= (int16_t)bit;
                                b.drop(1);
                            }
                            continue;
                        }
                        const JHuff& ha = J.ac[sc[i].ta];
                        if (ah == 0) {
                            if (eobrun) {
                                --eobrun;
                                continue;
                            }
                            for (int k = ss; k <= se;) {
                                b.fill();
                                const int rs = jdecode(b, ha);
                                if (rs < 0) return LLFE_E_INVALID;
                                const int r = rs >> 4, s2 = rs & 15;
                                if (s2 == 0) {
                                    if (r < 15) {
                                        eobrun = (1 << r) - 1;
                                        if (r) {
                                            eobrun += (int)b.peek(r);
                                            b.drop(r);
                                        }
                                        break;
                                    }
                                    k += 16;
                                } else {
                                    k += r;
                                    if (k > se) return LLFE_E_INVALID;
                                    blk[JZIGZAG[k]] = (int16_t)(jextend((int)b.peek(s2), s2) * bit);
                                    b.drop(s2);
                                    ++k;
                                }
                            }
                            continue;
                        }
                        // refinement of AC coefficients: a correction bit for every coefficient that is already non-zero,
                        // new +-1 coefficients (at this bit position) placed by zero runs over the still-zero ones
                        auto refine = [&](int16_t* p) {
                            b.fill();
                            if (b.peek(1) && (*p & bit) == 0) *p = (int16_t)(*p > 0 ? *p + bit : *p - bit);
                            b.drop(1);
                        };
                        if (eobrun) {
                            --eobrun;
                            for (int k = ss; k <= se; ++k)
                                if (blk[JZIGZAG[k]] != 0) refine(blk + JZIGZAG[k]);
                            continue;
                        }
                        for (int k = ss; k <= se;) {
                            b.fill();
                            const int rs = jdecode(b, ha);
                            if (rs < 0) return LLFE_E_INVALID;
                            int r = rs >> 4;
                            const int s2 = rs & 15;
                            int val = 0;
                            if (s2 == 0) {
                                if (r < 15) {
                                    eobrun = (1 << r) - 1;
                                    if (r) {
                                        eobrun += (int)b.peek(r);
                                        b.drop(r);
                                    }
                                    r = 64;       // the rest of the band only gets correction bits
                                }
                            } else {
                                if (s2 != 1) return LLFE_E_INVALID;
                                val = b.peek(1) ? bit : -bit;
                                b.drop(1);
                            }
                            while (k <= se) {
                                int16_t* p = blk + JZIGZAG[k++];
                                if (*p != 0) {
                                    refine(p);
                                } else {
                                    if (r == 0) {
                                        if (s2) *p = (int16_t)val;
                                        break;
                                    }
                                    --r;
                                }
                            }
                        }
                    }
            }
        }
        if (b.cnt < b.fake) return LLFE_E_INVALID;
        // ---- to the next marker: the reader stands on its FF if it has seen it, else search
        const uint8_t* q = b.p;
        if (!b.marker)
            while (q + 1 < buf + len && !(q[0] == 0xFF && q[1] != 0 && !(q[1] >= 0xD0 && q[1] <= 0xD7))) ++q;
        pos = (size_t)(q - buf);
        for (;;) {                    // tables may change between scans
            if (pos + 2 > len || buf[pos] != 0xFF) return LLFE_E_INVALID;
            while (pos < len && buf[pos] == 0xFF) ++pos;
            if (pos >= len) return LLFE_E_INVALID;
            const int m = buf[pos++];
            if (m == 0xD9) {
                for (int c = 0; c < J.ncomp; ++c)
                    for (int k = 0; k < 64; ++k)
                        if (prec[c][k] != 0) return LLFE_E_UNSUPPORTED;   // incomplete progression: libjpeg smooths such files
                return LLFE_OK;
            }
            if (pos + 2 > len) return LLFE_E_INVALID;
            const size_t sn = ((size_t)buf[pos] << 8) | buf[pos + 1];
            if (sn < 2 || pos + sn > len) return LLFE_E_INVALID;
            if (m == 0xDA) break;     // pos is at the length field of the next scan header
            const uint8_t* sg = buf + pos + 2;
            const size_t sl = sn - 2;
            if (m == 0xC4) {
                size_t o = 0;
                while (o < sl) {
                    if (o + 17 > sl) return LLFE_E_INVALID;
                    const int tc = sg[o] >> 4, th = sg[o] & 15;
                    if (tc > 1 || th > 3) return LLFE_E_INVALID;
                    int nvv = 0;
                    for (int i = 0; i < 16; ++i) nvv += sg[o + 1 + i];
                    if (nvv > 256 || o + 17 + nvv > sl) return LLFE_E_INVALID;
                    if (!build_huff(sg + o + 1, sg + o + 17, nvv, tc ? &J.ac[th] : &J.dc[th])) return LLFE_E_INVALID;
                    o += 17 + nvv;
                }
            } else if (m == 0xDD) {
                if (sl < 2) return LLFE_E_INVALID;
                J.ri = (sg[0] << 8) | sg[1];
            } else if (m == 0xDB || (m >= 0xC0 && m <= 0xCF)) {
                return LLFE_E_UNSUPPORTED;     // new quantisation tables / frames between scans
            }
            pos += sn;
        }
    }
}

// ---- device side ------------------------------------------------------------------------------------------------------
struct JDevComp {
    int bw, bh, cw, ch, h, v;
    unsigned long long coef_off, plane_off;
    int tq;
};
struct JDev {
    int width, height, ncomp, hmax, vmax;
    JDevComp c[3];
    uint16_t q[3][64];       // quantisation table of each component, natural order
};

__device__ __forceinline__ void jidct8(int* x, int shift) {   // jidctint.c, one pass over x[0..7]
    const int z2 = x[2], z3 = x[6];
    int z1 = (z2 + z3) * 4433;
    const int t2 = z1 + z3 * (-15137), t3 = z1 + z2 * 6270;
    const int t0 = (x[0] + x[4]) << 13, t1 = (x[0] - x[4]) << 13;
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    int o0 = x[7], o1 = x[5], o2 = x[3], o3 = x[1];
    z1 = o0 + o3;
    int y2 = o1 + o2, y3 = o0 + o2, y4 = o1 + o3;
    const int z5 = (y3 + y4) * 9633;
    o0 *= 2446;
    o1 *= 16819;
    o2 *= 25172;
    o3 *= 12299;
    z1 *= -7373;
    y2 *= -20995;
    y3 = y3 * (-16069) + z5;
    y4 = y4 * (-3196) + z5;
    o0 += z1 + y3;
    o1 += y2 + y4;
    o2 += y2 + y3;
    o3 += z1 + y4;
    const int r = 1 << (shift - 1);
    x[0] = (t10 + o3 + r) >> shift;
    x[7] = (t10 - o3 + r) >> shift;
    x[1] = (t11 + o2 + r) >> shift;
    x[6] = (t11 - o2 + r) >> shift;
    x[2] = (t12 + o1 + r) >> shift;
    x[5] = (t12 - o1 + r) >> shift;
    x[3] = (t13 + o0 + r) >> shift;
    x[4] = (t13 - o0 + r) >> shift;
}

// a thread per 8 x 8 block: dequantise, two IDCT passes, range limit, 8 rows of 8 bytes into the component plane
__global__ void __launch_bounds__(128) k_jpeg_idct(const int16_t* __restrict__ coef, JDev J, int comp, int blk0, int blk1,
                                                   uint8_t* __restrict__ planes) {
    const JDevComp c = J.c[comp];
    const int blk = blk0 + blockIdx.x * 128 + threadIdx.x;     // blocks [blk0, blk1) of the component
    if (blk >= blk1) return;
    const int by = blk / c.bw, bx = blk - by * c.bw;
    const uint4* src = reinterpret_cast<const uint4*>(coef + c.coef_off + (size_t)blk * 64);
    int x[64];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 v = src[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[8 * i + 2 * j] = (int)(int16_t)(w[j] & 0xffffu) * (int)J.q[comp][8 * i + 2 * j];
            x[8 * i + 2 * j + 1] = (int)(int16_t)(w[j] >> 16) * (int)J.q[comp][8 * i + 2 * j + 1];
        }
    }
    // pass 1: columns
#pragma unroll
    for (int col = 0; col < 8; ++col) {
        int t[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) t[r] = x[8 * r + col];
        jidct8(t, 13 - 2);
#pragma unroll
        for (int r = 0; r < 8; ++r) x[8 * r + col] = t[r];
    }
    uint8_t* out = planes + c.plane_off + ((size_t)by * 8) * (c.bw * 8) + bx * 8;
#pragma unroll
    for (int row = 0; row < 8; ++row) {
        jidct8(x + 8 * row, 13 + 2 + 3);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            lo |= (uint32_t)min(max(x[8 * row + k] + 128, 0), 255) << (8 * k);
            hi |= (uint32_t)min(max(x[8 * row + 4 + k] + 128, 0), 255) << (8 * k);
        }
        *reinterpret_cast<uint2*>(out + (size_t)row * (c.bw * 8)) = make_uint2(lo, hi);
    }
}

// chroma sample at full resolution: libjpeg's fancy up-sampling (edges replicated)
__device__ __forceinline__ int jchroma(const uint8_t* __restrict__ p, const JDevComp& c, int hs, int vs, int x, int y) {
    const int stride = c.bw * 8;
    if (hs == 1 && vs == 1) return p[(size_t)y * stride + x];
    const int cx = x >> 1;
    const int xo = (x & 1) ? min(cx + 1, c.cw - 1) : max(cx - 1, 0);     // the farther column
    if (vs == 1) {   // h2v1: (3 * near + far + 1 or 2) >> 2
        const uint8_t* r = p + (size_t)y * stride;
        return (3 * r[cx] + r[xo] + ((x & 1) ? 2 : 1)) >> 2;
    }
    const int cy = y >> 1;
    const int yo = (y & 1) ? min(cy + 1, c.ch - 1) : max(cy - 1, 0);
    const uint8_t* r0 = p + (size_t)cy * stride;
    const uint8_t* r1 = p + (size_t)yo * stride;
    const int near = 3 * r0[cx] + r1[cx], far = 3 * r0[xo] + r1[xo];   // column sums of the two nearest rows
    return (3 * near + far + ((x & 1) ? 7 : 8)) >> 4;
}

__global__ void __launch_bounds__(256) k_jpeg_to_bgr(const uint8_t* __restrict__ planes, JDev J, int y0, uint8_t* __restrict__ bgr) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = y0 + blockIdx.y;
    if (x >= J.width) return;
    const int yv = planes[J.c[0].plane_off + (size_t)y * (J.c[0].bw * 8) + x];
    uint8_t* o = bgr + ((size_t)y * J.width + x) * 3;
    if (J.ncomp == 1) {
        o[0] = o[1] = o[2] = (uint8_t)yv;
        return;
    }
    const int cb = jchroma(planes + J.c[1].plane_off, J.c[1], J.hmax, J.vmax, x, y) - 128;
    const int cr = jchroma(planes + J.c[2].plane_off, J.c[2], J.hmax, J.vmax, x, y) - 128;
    // jdcolor.c: FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554, ONE_HALF = 32768
    const int r = yv + ((91881 * cr + 32768) >> 16);
    const int g = yv + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
    const int b = yv + ((116130 * cb + 32768) >> 16);
    o[0] = (uint8_t)min(max(b, 0), 255);
    o[1] = (uint8_t)min(max(g, 0), 255);
    o[2] = (uint8_t)min(max(r, 0), 255);
}

}  // namespace

// Host-only: is this a JPEG the device path takes, and how large is its image?  out[0] = width, out[1] = height.
extern "C" int llfe_jpeg_info(const uint8_t* buf, size_t len, int32_t* out) {
    if (!buf || !out) {
        llfe_set_error("llfe_jpeg_info: invalid argument: null pointer");
        return LLFE_E_INVALID;
    }
    JInfo* J = new JInfo;
    const int rc = jpeg_parse(buf, len, J);
    if (rc == LLFE_OK) {
        out[0] = J->width;
        out[1] = J->height;
    } else {
        llfe_set_error(rc == LLFE_E_UNSUPPORTED ? "llfe_jpeg_info: not a baseline JFIF file of the supported subset"
                                                : "llfe_jpeg_info: damaged JPEG header");
    }
    delete J;
    return rc;
}

// Host-only (tests): the quantised coefficients of the file, blocks [by][bx][64] in natural order, component after
// component (MCU-padded block grids); *count = int16 values written, nothing is written when cap is too small.
extern "C" int llfe_jpeg_coefficients(const uint8_t* buf, size_t len, int16_t* out, size_t cap, size_t* count) {
    if (!buf || !count || (!out && cap)) {
        llfe_set_error("llfe_jpeg_coefficients: invalid argument: null pointer");
        return LLFE_E_INVALID;
    }
    JInfo* J = new JInfo;
    int rc = jpeg_parse(buf, len, J);
    if (rc == LLFE_OK) {
        *count = J->coef_count;
        if (cap >= J->coef_count) rc = J->progressive ? jpeg_progressive_decode(*J, buf, len, out) : jpeg_entropy_decode(*J, out);
    }
    if (rc != LLFE_OK) llfe_set_error("llfe_jpeg_coefficients: unsupported or damaged file");
    delete J;
    return rc;
}

// One baseline JPEG -> BGR (h x w as reported by llfe_jpeg_info): entropy decoding on the calling thread into pinned memory
// (`pin`), IDCT / up-sampling / colour conversion on the device (`dev`); both buffers hold at least llfe_jpeg_stage_bytes.
int llfe_jpeg_decode_impl(llfe_ctx* ctx, const uint8_t* buf, size_t len, int h, int w, uint8_t* h_bgr, uint8_t* pin,
                          uint8_t* dev, size_t cap) {
    JInfo* J = new JInfo;
    struct Free {
        JInfo* j;
        ~Free() { delete j; }
    } fr{J};
    int rc = jpeg_parse(buf, len, J);
    if (rc != LLFE_OK) {
        llfe_set_error("llfe_jpeg_decode_host: file outside the supported subset or damaged header");
        return rc;
    }
    if (J->width != w || J->height != h) {
        llfe_set_error("llfe_jpeg_decode_host: size does not match the file (%d x %d)", J->width, J->height);
        return LLFE_E_INVALID;
    }
    const size_t coef_bytes = WsCarver::need(J->coef_count * 2), out = (size_t)h * w * 3;
    if (coef_bytes + WsCarver::need(J->plane_bytes) + WsCarver::need(out) > cap) {
        llfe_set_error("llfe_jpeg_decode_host: staging too small");
        return LLFE_E_INVALID;
    }
    int16_t* p_coef = reinterpret_cast<int16_t*>(pin);
    int16_t* d_coef = reinterpret_cast<int16_t*>(dev);
    uint8_t* d_planes = dev + coef_bytes;
    uint8_t* d_out = d_planes + WsCarver::need(J->plane_bytes);
    uint8_t* p_out = pin + coef_bytes;
    JDev D;
    D.width = w, D.height = h, D.ncomp = J->ncomp, D.hmax = J->hmax, D.vmax = J->vmax;
    for (int i = 0; i < J->ncomp; ++i) {
        const JComp& c = J->comp[i];
        D.c[i] = JDevComp{c.bw, c.bh, c.cw, c.ch, c.h, c.v, (unsigned long long)c.coef_off, (unsigned long long)c.plane_off, c.tq};
        memcpy(D.q[i], J->qt[c.tq], sizeof(D.q[i]));
    }
    // MCU rows [m0, m1) are final: copy their coefficients in, transform them, convert the pixel rows whose chroma
    // neighbours are final too, and bring those rows back
    int rows_out = 0;       // pixel rows converted so far
    auto ship = [&](int m0, int m1, bool last) -> int {
        for (int i = 0; i < J->ncomp; ++i) {
            const JComp& c = J->comp[i];
            const size_t b0 = (size_t)m0 * c.v * c.bw, b1 = (size_t)m1 * c.v * c.bw;
            LLFE_CUDA(cudaMemcpyAsync(d_coef + c.coef_off + b0 * 64, p_coef + c.coef_off + b0 * 64, (b1 - b0) * 128,
                                      cudaMemcpyHostToDevice, ctx->stream));
            LLFE_KERNEL(ctx, "k_jpeg_idct");
            k_jpeg_idct<<<ceil_div((int)(b1 - b0), 128), 128, 0, ctx->stream>>>(d_coef, D, i, (int)b0, (int)b1, d_planes);
            LLFE_LAUNCHED(ctx);
        }
        int y1 = last ? h : m1 * 8 * J->vmax - 8;      // the up-sampler looks one chroma row ahead
        if (y1 > h) y1 = h;
        if (y1 > rows_out) {
            LLFE_KERNEL(ctx, "k_jpeg_to_bgr");
            k_jpeg_to_bgr<<<dim3(ceil_div(w, 256), y1 - rows_out), 256, 0, ctx->stream>>>(d_planes, D, rows_out, d_out);
            LLFE_LAUNCHED(ctx);
            const size_t o0 = (size_t)rows_out * w * 3, o1 = (size_t)y1 * w * 3;
            LLFE_CUDA(cudaMemcpyAsync(p_out + o0, d_out + o0, o1 - o0, cudaMemcpyDeviceToHost, ctx->stream));
            rows_out = y1;
        }
        return LLFE_OK;
    };
    const int bands = (!J->progressive && J->coef_count * 2 >= (size_t(1) << 20) && J->mcuy >= 16) ? 8 : 1;
    if (bands == 1) {
        rc = J->progressive ? jpeg_progressive_decode(*J, buf, len, p_coef) : jpeg_entropy_decode(*J, p_coef);
        if (rc != LLFE_OK) {
            llfe_set_error("llfe_jpeg_decode_host: damaged entropy-coded data");
            return rc;
        }
        LLFE_TRY(ship(0, J->mcuy, true));
        LLFE_CUDA(cudaStreamSynchronize(ctx->stream));
        memcpy(h_bgr, p_out, out);
        return LLFE_OK;
    }
    // The entropy decoding runs on a helper thread that reports every finished MCU row; this thread ships band after band
    // (copy in, IDCT, conversion, copy back) and moves finished rows into the caller's buffer meanwhile, so that only the
    // last band's share of the copies follows the decoder.
    std::atomic<int> rows_done{0};
    std::atomic<int> done{0};
    int ent_rc = LLFE_OK;
    std::thread worker([&] {
        ent_rc = jpeg_entropy_decode(*J, p_coef, &rows_done);
        done.store(1, std::memory_order_release);
    });
    const int per = ceil_div(J->mcuy, bands);
    cudaEvent_t ev = nullptr;
    int copied = 0, pending_rows = 0;
    rc = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess ? LLFE_OK : LLFE_E_CUDA;
    for (int m0 = 0; m0 < J->mcuy && rc == LLFE_OK; m0 += per) {
        const int m1 = m0 + per < J->mcuy ? m0 + per : J->mcuy;
        while (rows_done.load(std::memory_order_acquire) < m1 && !done.load(std::memory_order_acquire)) std::this_thread::yield();
        if (rows_done.load(std::memory_order_acquire) < m1) break;      // the decoder gave up
        if (pending_rows > copied) {       // rows of the previous band are back by now or soon: hand them over
            cudaEventSynchronize(ev);
            memcpy(h_bgr + (size_t)copied * w * 3, p_out + (size_t)copied * w * 3, (size_t)(pending_rows - copied) * w * 3);
            copied = pending_rows;
        }
        rc = ship(m0, m1, m1 == J->mcuy);
        if (rc == LLFE_OK) {
            cudaEventRecord(ev, ctx->stream);
            pending_rows = rows_out;
        }
    }
    worker.join();
    cudaStreamSynchronize(ctx->stream);
    if (ev) cudaEventDestroy(ev);
    if (rc != LLFE_OK) return rc;
    if (ent_rc != LLFE_OK || rows_out != h) {
        llfe_set_error("llfe_jpeg_decode_host: damaged entropy-coded data");
        return ent_rc != LLFE_OK ? ent_rc : LLFE_E_INVALID;
    }
    memcpy(h_bgr + (size_t)copied * w * 3, p_out + (size_t)copied * w * 3, (size_t)(h - copied) * w * 3);
    return LLFE_OK;
}

// staging need of llfe_jpeg_decode_impl for a w x h image (4:4:4 is the largest: three full planes of int16 + u8)
size_t llfe_jpeg_stage_bytes(int h, int w) {
    const size_t bw = (size_t)(w + 15) / 16 * 2, bh = (size_t)(h + 15) / 16 * 2;
    const size_t blocks = 3 * bw * bh;
    return WsCarver::need(blocks * 128) + WsCarver::need(blocks * 64 + 3 * 256) + WsCarver::need((size_t)h * w * 3);
}
