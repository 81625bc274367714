// Host side of the PNG decode row: RFC 1950 / 1951 inflate of the IDAT stream.
//
// A deflate stream is one serial bit-level decode (every code's position depends on all codes before it), so it runs on a
// host core.  cv2.imdecode spends ~80 % of a 1080p PNG decode inside zlib 1.2.11's inflate.  This decoder keeps a 64-bit
// bit buffer refilled with one unaligned load, resolves most codes with a single look-up in an 11-bit (literal / length)
// or 8-bit (distance) root table, decodes up to three literals per refill and copies matches eight bytes at a time; on
// the noisy design images of the benchmark (one literal per byte) it is bound by the look-up -> shift -> look-up
// dependency chain and runs 1.1-1.4x zlib (measured: 30 vs 42 ms per 1080p stream in the build container, 24.7 vs
// 27.7 ms on the GPU box); long-match data gains more.  Its other job is to write straight into the pinned staging
// buffer of the context, from where the scanlines go to the device (k_png.cu) without a host copy.
//
// Error behaviour follows zlib's inflate: over-subscribed or incomplete code sets, a missing end-of-block code, invalid
// symbols, distances beyond the start of the output, a stored block whose LEN / NLEN disagree, a truncated stream and an
// Adler-32 mismatch are all failures (the caller then leaves the file to cv2.imdecode).  Output beyond `out_cap` is not an
// error: libpng stops reading once the image is complete ("too much image data" is a warning).
#include <string.h>

#include <atomic>

#include "llfe_common.cuh"

namespace {

struct Ent {
    uint16_t val;
    uint8_t op;     // 0 literal; 16|extra length / distance base; 32|n link to a sub-table of 2^n entries; 64 end of block; 128 invalid
    uint8_t bits;   // bits this entry consumes
};
constexpr uint8_t OP_LIT = 0, OP_BASE = 16, OP_SUB = 32, OP_END = 64, OP_BAD = 128;
constexpr int LBITS = 11, DBITS = 8, PBITS = 7;
constexpr int LCAP = (1 << LBITS) + 288 * 16, DCAP = (1 << DBITS) + 32 * 128, PCAP = 1 << PBITS;

const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

enum Kind { LITLEN, DIST, PRECODE };

inline Ent make_ent(Kind kind, int sym, int bits) {
    Ent e;
    e.bits = (uint8_t)bits;
    if (kind == PRECODE) {
        e.val = (uint16_t)sym, e.op = OP_LIT;
    } else if (kind == LITLEN) {
        if (sym < 256) e.val = (uint16_t)sym, e.op = OP_LIT;
        else if (sym == 256) e.val = 0, e.op = OP_END;
        else if (sym < 286) e.val = LBASE[sym - 257], e.op = (uint8_t)(OP_BASE | LEXT[sym - 257]);
        else e.val = 0, e.op = OP_BAD;
    } else {
        if (sym < 30) e.val = DBASE[sym], e.op = (uint8_t)(OP_BASE | DEXT[sym]);
        else e.val = 0, e.op = OP_BAD;
    }
    return e;
}

inline uint32_t bitrev(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}

// canonical Huffman code of `n` symbols with lengths lens[] (0 = unused) -> root table of 2^root entries + sub-tables
bool build_table(const uint8_t* lens, int n, Kind kind, int root, Ent* table, int cap) {
    int count[16] = {0};
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    int maxlen = 15;
    while (maxlen > 0 && count[maxlen] == 0) --maxlen;
    const Ent bad = {0, OP_BAD, 1};
    for (int i = 0; i < (1 << root); ++i) table[i] = bad;
    if (maxlen == 0) return kind != PRECODE;   // no codes at all: every look-up is invalid (zlib accepts the set itself)
    int left = 1;
    for (int len = 1; len <= 15; ++len) {
        left <<= 1;
        left -= count[len];
        if (left < 0) return false;             // over-subscribed
    }
    if (left > 0 && (kind == PRECODE || maxlen != 1)) return false;   // incomplete (zlib allows one single-bit code)
    uint32_t next_code[16];
    uint32_t code = 0;
    count[0] = 0;
    for (int len = 1; len <= 15; ++len) {
        code = (code + count[len - 1]) << 1;
        next_code[len] = code;
    }
    const int sub_bits = maxlen > root ? maxlen - root : 0;
    int next = 1 << root;
    for (int sym = 0; sym < n; ++sym) {
        const int len = lens[sym];
        if (!len) continue;
        const uint32_t rev = bitrev(next_code[len]++, len);
        if (len <= root) {
            const Ent e = make_ent(kind, sym, len);
            for (uint32_t i = rev; i < (1u << root); i += 1u << len) table[i] = e;
        } else {
            const uint32_t prefix = rev & ((1u << root) - 1);
            if (!(table[prefix].op & OP_SUB)) {
                if (next + (1 << sub_bits) > cap) return false;
                table[prefix].val = (uint16_t)next, table[prefix].op = (uint8_t)(OP_SUB | sub_bits), table[prefix].bits = (uint8_t)root;
                for (int i = 0; i < (1 << sub_bits); ++i) table[next + i] = bad;
                next += 1 << sub_bits;
            }
            Ent* sub = table + table[prefix].val;
            const Ent e = make_ent(kind, sym, len - root);
            for (uint32_t i = rev >> root; i < (1u << sub_bits); i += 1u << (len - root)) sub[i] = e;
        }
    }
    return true;
}

struct Tables {
    Ent lit[LCAP];
    Ent dist[DCAP];
};

struct Fixed {
    Tables t;
    Fixed() {
        uint8_t l[288];
        for (int i = 0; i < 144; ++i) l[i] = 8;
        for (int i = 144; i < 256; ++i) l[i] = 9;
        for (int i = 256; i < 280; ++i) l[i] = 7;
        for (int i = 280; i < 288; ++i) l[i] = 8;
        build_table(l, 288, LITLEN, LBITS, t.lit, LCAP);
        uint8_t d[32];
        for (int i = 0; i < 32; ++i) d[i] = 5;
        build_table(d, 32, DIST, DBITS, t.dist, DCAP);
    }
};

inline uint64_t load64(const uint8_t* p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;   // little endian hosts only (x86-64 / aarch64)
}

struct Stream {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t buf = 0;
    int cnt = 0;   // valid bits in buf
    // at least 56 valid bits, or everything that is left of the input
    inline void refill() {
        if (end - p >= 8) {
            buf |= load64(p) << cnt;
            p += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56 && p < end) {
                buf |= (uint64_t)*p++ << cnt;
                cnt += 8;
            }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    inline void drop(int n) {
        buf >>= n;
        cnt -= n;
    }
};

enum { INF_OK = 0, INF_FULL = 1, INF_ERR = -1 };

// one Huffman-coded block: symbols until end-of-block.  Returns INF_OK at end of block, INF_FULL when a symbol would write
// past out_end, INF_ERR on invalid data / truncated input.
int decode_block(Stream& s_ref, const Ent* lt, const Ent* dt, uint8_t* out_begin, uint8_t*& out_ref, uint8_t* out_end) {
    Stream s = s_ref;          // a local copy: byte stores to `out` may alias anything reachable through a reference
    uint8_t* out = out_ref;
    int rc = INF_ERR;
    for (;;) {
        s.refill();
        int lits = 0;
    again:
        Ent e = lt[s.buf & ((1u << LBITS) - 1)];
        if (e.op & OP_SUB) {
            if (s.cnt < e.bits) break;
            s.drop(e.bits);
            e = lt[e.val + s.peek(e.op & 15)];
        }
        if (e.bits > s.cnt) break;              // truncated input (or an invalid entry at the very end)
        s.drop(e.bits);
        if (e.op == OP_LIT) {
            if (out >= out_end) {
                rc = INF_FULL;
                break;
            }
            *out++ = (uint8_t)e.val;
            if (++lits < 3 && s.cnt >= 15) goto again;   // up to three 15-bit codes per refill
            continue;
        }
        if (e.op & OP_BASE) {
            const int lext = e.op & 15;
            if (s.cnt < 33) s.refill();          // length extra (5) + distance code (15) + distance extra (13)
            if (s.cnt < lext) break;
            uint32_t len = e.val + s.peek(lext);
            s.drop(lext);
            Ent d = dt[s.buf & ((1u << DBITS) - 1)];
            if (d.op & OP_SUB) {
                if (s.cnt < d.bits) break;
                s.drop(d.bits);
                d = dt[d.val + s.peek(d.op & 15)];
            }
            if (!(d.op & OP_BASE) || d.op == OP_BAD) break;     // invalid distance code
            const int dext = d.op & 15;
            if (d.bits + dext > s.cnt) break;
            s.drop(d.bits);
            const uint32_t dist = d.val + s.peek(dext);
            s.drop(dext);
            if (dist > (size_t)(out - out_begin)) break;        // distance too far back
            if (len > (size_t)(out_end - out)) {
                // the image is complete before the match is: fill what fits and stop (libpng: too much image data)
                const uint8_t* src = out - dist;
                while (out < out_end) *out++ = *src++;
                rc = INF_FULL;
                break;
            }
            const uint8_t* src = out - dist;
            if (dist >= 8 && (size_t)(out_end - out) >= len + 8) {
                uint8_t* o = out;
                const uint8_t* e8 = out + len;
                do {
                    memcpy(o, src, 8);
                    o += 8, src += 8;
                } while (o < e8);
            } else if (dist == 1) {
                memset(out, *src, len);
            } else {
                for (uint32_t i = 0; i < len; ++i) out[i] = src[i];
            }
            out += len;
            continue;
        }
        if (e.op == OP_END) rc = INF_OK;
        break;
    }
    out_ref = out;
    s_ref = s;
    return rc;
}

int inflate_raw(Stream& s, uint8_t* out, size_t out_cap, size_t* out_len, bool* finished, std::atomic<size_t>* progress) {
    static const Fixed fixed;
    static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    Tables* dyn = nullptr;
    uint8_t* o = out;
    uint8_t* const out_end = out + out_cap;
    int rc = INF_ERR;
    *finished = false;
    for (;;) {
        s.refill();
        if (s.cnt < 3) break;
        const int last = s.peek(1);
        s.drop(1);
        const int type = s.peek(2);
        s.drop(2);
        if (type == 0) {
            s.drop(s.cnt & 7);                       // to the next byte boundary
            if (s.cnt < 32) s.refill();
            if (s.cnt < 32) break;
            const uint32_t len = s.peek(16);
            s.drop(16);
            const uint32_t nlen = s.peek(16);
            s.drop(16);
            if ((len ^ 0xffffu) != nlen) break;
            // give the whole bytes of the bit buffer back to the input
            s.p -= s.cnt >> 3;
            s.buf = 0, s.cnt = 0;
            if ((size_t)(s.end - s.p) < len) break;
            const size_t room = (size_t)(out_end - o);
            const size_t take = len < room ? len : room;
            memcpy(o, s.p, take);
            o += take;
            s.p += len;
            if (take < len) {
                rc = INF_FULL;
                break;
            }
        } else if (type == 1 || type == 2) {
            const Ent *lt, *dt;
            if (type == 1) {
                lt = fixed.t.lit, dt = fixed.t.dist;
            } else {
                if (s.cnt < 14) break;
                const int nlen = s.peek(5) + 257;
                s.drop(5);
                const int ndist = s.peek(5) + 1;
                s.drop(5);
                const int ncode = s.peek(4) + 4;
                s.drop(4);
                if (nlen > 286 || ndist > 30) break;
                uint8_t lens[320];
                memset(lens, 0, 19);
                bool ok = true;
                for (int i = 0; i < ncode; ++i) {
                    if (s.cnt < 3) s.refill();
                    if (s.cnt < 3) {
                        ok = false;
                        break;
                    }
                    lens[ORDER[i]] = (uint8_t)s.peek(3);
                    s.drop(3);
                }
                Ent pre[PCAP];
                if (!ok || !build_table(lens, 19, PRECODE, PBITS, pre, PCAP)) break;
                int have = 0;
                const int total = nlen + ndist;
                while (have < total) {
                    s.refill();
                    const Ent e = pre[s.buf & ((1u << PBITS) - 1)];
                    if (e.op != OP_LIT || e.bits > s.cnt) {
                        ok = false;
                        break;
                    }
                    const int sym = e.val;
                    int rep, val;
                    if (sym < 16) {
                        s.drop(e.bits);
                        lens[have++] = (uint8_t)sym;
                        continue;
                    }
                    const int xb = sym == 16 ? 2 : sym == 17 ? 3 : 7;
                    if (e.bits + xb > s.cnt) {
                        ok = false;
                        break;
                    }
                    s.drop(e.bits);
                    if (sym == 16) {
                        if (have == 0) {
                            ok = false;
                            break;
                        }
                        val = lens[have - 1];
                        rep = 3 + s.peek(2);
                    } else {
                        val = 0;
                        rep = sym == 17 ? 3 + s.peek(3) : 11 + s.peek(7);
                    }
                    s.drop(xb);
                    if (have + rep > total) {
                        ok = false;
                        break;
                    }
                    while (rep--) lens[have++] = (uint8_t)val;
                }
                if (!ok || lens[256] == 0) break;            // missing end-of-block code
                if (!dyn) dyn = new Tables;
                uint8_t dl[32];
                memcpy(dl, lens + nlen, ndist);
                if (!build_table(lens, nlen, LITLEN, LBITS, dyn->lit, LCAP) || !build_table(dl, ndist, DIST, DBITS, dyn->dist, DCAP))
                    break;
                lt = dyn->lit, dt = dyn->dist;
            }
            const int r = decode_block(s, lt, dt, out, o, out_end);
            if (r != INF_OK) {
                rc = r;
                break;
            }
        } else {
            break;   // reserved block type
        }
        if (progress) progress->store((size_t)(o - out), std::memory_order_release);   // a consumer may take the bytes so far
        if (last) {
            rc = INF_OK;
            *finished = true;
            break;
        }
    }
    delete dyn;
    *out_len = (size_t)(o - out);
    return rc;
}

uint32_t adler32(const uint8_t* p, size_t n) {
    uint32_t a = 1, b = 0;
    while (n) {
        const size_t m = n < 5552 ? n : 5552;
        // b += m * a + sum (m - i) p[i];  a += sum p[i]   (reductions without a loop-carried chain: vectorisable)
        uint32_t s1 = 0, s2 = 0;
        for (size_t i = 0; i < m; ++i) {
            s1 += p[i];
            s2 += (uint32_t)(m - i) * p[i];
        }
        b = (b + (uint32_t)((uint64_t)m * a % 65521u) + s2 % 65521u) % 65521u;
        a = (a + s1) % 65521u;
        p += m, n -= m;
    }
    return (b << 16) | a;
}

}  // namespace

// zlib-wrapped deflate stream -> out (at most out_cap bytes).  *out_len = bytes written.  LLFE_OK when the stream is valid
// as far as it was needed: it ended (then the Adler-32 trailer must match) or the output filled up first.  `progress`
// (optional) is advanced to the number of finished output bytes after every deflate block, so that another thread can
// ship the front of the output while the rest is still being decoded (bytes below the mark never change again).
int llfe_inflate_zlib_progress(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len,
                               std::atomic<size_t>* progress) {
    if (!in || !out || !out_len) {
        llfe_set_error("llfe_inflate_zlib: invalid argument: null pointer");
        return LLFE_E_INVALID;
    }
    *out_len = 0;
    if (in_len < 2 || (in[0] & 15) != 8 || (in[0] >> 4) > 7 || (in[1] & 32) || ((in[0] << 8) | in[1]) % 31) {
        llfe_set_error("llfe_inflate_zlib: incorrect header check");
        return LLFE_E_INVALID;
    }
    Stream s;
    s.p = in + 2, s.end = in + in_len;
    bool finished = false;
    const int rc = inflate_raw(s, out, out_cap, out_len, &finished, progress);
    if (rc == INF_ERR) {
        llfe_set_error("llfe_inflate_zlib: invalid or truncated deflate stream");
        return LLFE_E_INVALID;
    }
    if (finished) {
        s.drop(s.cnt & 7);
        s.p -= s.cnt >> 3;
        if (s.end - s.p < 4) {
            llfe_set_error("llfe_inflate_zlib: truncated stream (no check value)");
            return LLFE_E_INVALID;
        }
        const uint32_t want = ((uint32_t)s.p[0] << 24) | ((uint32_t)s.p[1] << 16) | ((uint32_t)s.p[2] << 8) | s.p[3];
        if (adler32(out, *out_len) != want) {
            llfe_set_error("llfe_inflate_zlib: incorrect data check");
            return LLFE_E_INVALID;
        }
    }
    if (progress) progress->store(*out_len, std::memory_order_release);
    return LLFE_OK;
}

extern "C" int llfe_inflate_zlib(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len) {
    return llfe_inflate_zlib_progress(in, in_len, out, out_cap, out_len, nullptr);
}
