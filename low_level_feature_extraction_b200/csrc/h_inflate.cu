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
#include <mutex>
#include <thread>
#include <vector>

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

// ---- speculative workers (one stream decoded by several threads) ---------------------------------------------------
// A worker starts somewhere in the middle of the compressed data: it looks, bit by bit, for the header of a dynamic
// block that parses (complete code sets, an end-of-block code) and decodes from there without knowing the 32 KB of
// output in front of it -- into 16-bit symbols: a literal, or 0x8000 | i for "byte i of the unknown window", which match
// copies carry along like literals.  The decoder in front of it stops when it arrives, at a block boundary, at exactly
// the bit the worker started from; that proves the start was a real block start, fixes the worker's place in the output,
// and the symbols are turned into bytes (the window is the finished output in front of them).  A start that is never
// arrived at, a worker that finds none, or one that fails, only costs time: the decoder in front simply goes on.
struct SpecWorker {
    long long scan_from = 0;                   // first bit (relative to the deflate data) this worker may start at
    std::atomic<long long> start_bit{-1};      // published once the first block from there has decoded
    std::atomic<int> cancel{0};
    std::atomic<int> have_off{0};              // the decoder in front has arrived: `off` is this worker's place in the output
    std::atomic<int> resolved{0};              // 1 = bytes final (or failed: see ok)
    size_t off = 0;
    std::vector<uint16_t> sym;
    size_t len = 0;                            // symbols decoded
    long long end_bit = 0;                     // where it stopped (a block boundary, or behind the final block)
    int next = -1;                             // the worker it handed over to
    int pred = 0;                              // the decoder that arrived at this worker's start (0 = the plain one)
    bool finished = false;                     // saw the final block
    bool full = false;                         // stopped because the output cannot be longer than this
    bool ok = false;
    uint32_t adler = 1;                        // of its bytes
    size_t bytes = 0;                          // bytes it put into the output
};

struct SpecCtl {
    const uint8_t* base;       // first byte of the deflate data
    const uint8_t* end;
    uint8_t* out;
    size_t out_cap;
    int n = 0;                 // workers (index 0 is the plain decoder at the front and has no SpecWorker duties)
    SpecWorker* w = nullptr;
    std::atomic<int> all_done{0};
};

// The symbol buffers are kept between calls: a fresh multi-megabyte allocation per worker and call is served by mmap,
// and the page faults of several threads filling new mappings at once cost more than the decode itself.
struct SymPool {
    std::mutex m;
    std::vector<std::vector<uint16_t>> idle;
    void take(std::vector<uint16_t>& v) {
        std::lock_guard<std::mutex> g(m);
        if (!idle.empty()) {
            v.swap(idle.back());
            idle.pop_back();
        }
    }
    void give(std::vector<uint16_t>& v) {
        std::lock_guard<std::mutex> g(m);
        if (idle.size() < 8 && v.capacity() <= (size_t(32) << 20)) {   // (64 MB of symbols at most are kept per buffer)
            idle.emplace_back();
            idle.back().swap(v);
        }
    }
};
SymPool& sym_pool() {
    static SymPool p;
    return p;
}

inline long long bit_pos(const Stream& s, const uint8_t* base) { return (long long)(s.p - base) * 8 - s.cnt; }

inline void seek_bit(Stream& s, const uint8_t* base, const uint8_t* end, long long bit) {
    s.p = base + (bit >> 3), s.end = end, s.buf = 0, s.cnt = 0;
    s.refill();
    s.drop((int)(bit & 7));
}

// at a block boundary of decoder `self`: the worker that starts exactly here, or -1.  Workers whose start (or whose
// whole search range) lies behind are cancelled: nobody can arrive at them any more.
// Only a decoder whose own position is proven (the plain decoder at the front, a worker that has been arrived at) may
// cancel: a worker on a false start decodes garbage and must not take real workers down with it.
int spec_handover(SpecCtl* c, int self, long long bit, bool may_cancel) {
    for (int j = self + 1; j < c->n; ++j) {
        SpecWorker& w = c->w[j];
        if (w.cancel.load(std::memory_order_relaxed)) continue;
        const long long st = w.start_bit.load(std::memory_order_acquire);
        if (st == bit) return j;
        if (may_cancel && (st >= 0 ? st < bit : w.scan_from < bit)) w.cancel.store(1, std::memory_order_release);
    }
    return -1;
}

enum { SYM_ROOM = 264 };   // free symbols a step of decode_block_sym may need (three literals or one match)

// decode_block with 16-bit symbols and an unknown window: INF_OK at end of block, INF_FULL when fewer than SYM_ROOM
// symbols are free (nothing consumed: call again with more room), INF_ERR on invalid data / truncated input
int decode_block_sym(Stream& s_ref, const Ent* lt, const Ent* dt, uint16_t* out_begin, uint16_t*& out_ref, uint16_t* out_end) {
    Stream s = s_ref;
    uint16_t* out = out_ref;
    int rc = INF_ERR;
    for (;;) {
        if (out_end - out < SYM_ROOM) {
            rc = INF_FULL;
            break;
        }
        s.refill();
        int lits = 0;
    again:
        Ent e = lt[s.buf & ((1u << LBITS) - 1)];
        if (e.op & OP_SUB) {
            if (s.cnt < e.bits) break;
            s.drop(e.bits);
            e = lt[e.val + s.peek(e.op & 15)];
        }
        if (e.bits > s.cnt) break;
        s.drop(e.bits);
        if (e.op == OP_LIT) {
            *out++ = e.val;
            if (++lits < 3 && s.cnt >= 15) goto again;
            continue;
        }
        if (e.op & OP_BASE) {
            const int lext = e.op & 15;
            if (s.cnt < 33) s.refill();
            if (s.cnt < lext) break;
            const uint32_t len = e.val + s.peek(lext);
            s.drop(lext);
            Ent d = dt[s.buf & ((1u << DBITS) - 1)];
            if (d.op & OP_SUB) {
                if (s.cnt < d.bits) break;
                s.drop(d.bits);
                d = dt[d.val + s.peek(d.op & 15)];
            }
            if (!(d.op & OP_BASE) || d.op == OP_BAD) break;
            const int dext = d.op & 15;
            if (d.bits + dext > s.cnt) break;
            s.drop(d.bits);
            const uint32_t dist = d.val + s.peek(dext);
            s.drop(dext);
            const long long pos = out - out_begin;
            if (dist <= pos) {
                const uint16_t* src = out - dist;
                if (dist >= len) memcpy(out, src, 2 * (size_t)len);
                else
                    for (uint32_t i = 0; i < len; ++i) out[i] = src[i];
            } else {
                for (uint32_t i = 0; i < len; ++i) {   // (the first 32 KB only) reaches into the unknown window
                    const long long si = pos + i - dist;
                    out[i] = si >= 0 ? out_begin[si] : (uint16_t)(0x8000 | (32768 + si));
                }
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

// the code-length section of a dynamic block (RFC 1951 3.2.7), the stream positioned behind the three header bits
bool read_dynamic_tables(Stream& s, Tables* dyn) {
    static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    if (s.cnt < 14) return false;
    const int nlen = s.peek(5) + 257;
    s.drop(5);
    const int ndist = s.peek(5) + 1;
    s.drop(5);
    const int ncode = s.peek(4) + 4;
    s.drop(4);
    if (nlen > 286 || ndist > 30) return false;
    uint8_t lens[320];
    memset(lens, 0, 19);
    for (int i = 0; i < ncode; ++i) {
        if (s.cnt < 3) s.refill();
        if (s.cnt < 3) return false;
        lens[ORDER[i]] = (uint8_t)s.peek(3);
        s.drop(3);
    }
    Ent pre[PCAP];
    if (!build_table(lens, 19, PRECODE, PBITS, pre, PCAP)) return false;
    int have = 0;
    const int total = nlen + ndist;
    while (have < total) {
        s.refill();
        const Ent e = pre[s.buf & ((1u << PBITS) - 1)];
        if (e.op != OP_LIT || e.bits > s.cnt) return false;
        const int sym = e.val;
        int rep, val;
        if (sym < 16) {
            s.drop(e.bits);
            lens[have++] = (uint8_t)sym;
            continue;
        }
        const int xb = sym == 16 ? 2 : sym == 17 ? 3 : 7;
        if (e.bits + xb > s.cnt) return false;
        s.drop(e.bits);
        if (sym == 16) {
            if (have == 0) return false;
            val = lens[have - 1];
            rep = 3 + s.peek(2);
        } else {
            val = 0;
            rep = sym == 17 ? 3 + s.peek(3) : 11 + s.peek(7);
        }
        s.drop(xb);
        if (have + rep > total) return false;
        while (rep--) lens[have++] = (uint8_t)val;
    }
    if (lens[256] == 0) return false;            // missing end-of-block code
    uint8_t dl[32];
    memcpy(dl, lens + nlen, ndist);
    return build_table(lens, nlen, LITLEN, LBITS, dyn->lit, LCAP) && build_table(dl, ndist, DIST, DBITS, dyn->dist, DCAP);
}

const Fixed& fixed_tables() {
    static const Fixed fixed;
    return fixed;
}

enum { INF_HANDED = 2 };   // inflate_raw: stopped at the block boundary where worker *handed_to starts

// the blocks from the position of `s` on, output from out + start_off (matches may reach back to out).  With `ctl`, the
// decoder looks at every block boundary for a speculative worker that started exactly there and stops if there is one.
int inflate_raw(Stream& s, uint8_t* out, size_t start_off, size_t out_cap, size_t* out_len, bool* finished,
                std::atomic<size_t>* progress, SpecCtl* ctl = nullptr, int* handed_to = nullptr) {
    const Fixed& fixed = fixed_tables();
    Tables* dyn = nullptr;
    uint8_t* o = out + start_off;
    uint8_t* const out_end = out + out_cap;
    int rc = INF_ERR;
    *finished = false;
    for (;;) {
        s.refill();
        if (ctl) {
            const int j = spec_handover(ctl, 0, bit_pos(s, ctl->base), true);
            if (j >= 0) {
                *handed_to = j;
                rc = INF_HANDED;
                break;
            }
        }
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
                if (!dyn) dyn = new Tables;
                if (!read_dynamic_tables(s, dyn)) break;
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


// zlib's adler32_combine: the check value of A followed by B from those of A and B and the length of B
uint32_t adler32_combine(uint32_t a1, uint32_t a2, size_t len2) {
    const uint32_t BASE = 65521u;
    const uint32_t rem = (uint32_t)(len2 % BASE);
    uint32_t sum1 = a1 & 0xffff;
    uint32_t sum2 = (uint32_t)(((uint64_t)rem * sum1) % BASE);
    sum1 += (a2 & 0xffff) + BASE - 1;
    sum2 += ((a1 >> 16) & 0xffff) + ((a2 >> 16) & 0xffff) + BASE - rem;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum2 >= (BASE << 1)) sum2 -= (BASE << 1);
    if (sum2 >= BASE) sum2 -= BASE;
    return sum1 | (sum2 << 16);
}

uint32_t adler32(const uint8_t* p, size_t n);

// worker k: find a start, decode until handed over / end of stream / full / failure, then -- once the decoder in front
// has arrived -- turn the symbols into bytes
void spec_run(SpecCtl* c, int k) {
    SpecWorker& w = c->w[k];
    struct Back {
        std::vector<uint16_t>& v;
        ~Back() { sym_pool().give(v); }
    } back{w.sym};
    const Fixed& fixed = fixed_tables();
    Tables* dyn = new Tables;
    const long long total_bits = (long long)(c->end - c->base) * 8;
    const long long scan_limit = w.scan_from + (1ll << 21);   // 256 KB of compressed data without a block start: give up
    // a worker cannot contribute more than the whole output: beyond that the output is full wherever its place is
    const size_t max_syms = c->out_cap + SYM_ROOM + 1;
    sym_pool().take(w.sym);
    {
        size_t cap = c->out_cap / (size_t)c->n + c->out_cap / 8 + 65536;
        if (cap > max_syms) cap = max_syms;
        if (w.sym.size() < cap) w.sym.resize(cap);
    }
    auto stop_asked = [&]() { return w.cancel.load(std::memory_order_relaxed) || c->all_done.load(std::memory_order_relaxed); };
    // room for `need` more symbols behind `have`; false = the output is full before that
    auto room = [&](size_t have, size_t need) -> bool {
        if (have + need <= w.sym.size()) return true;
        if (have + need > max_syms) return false;
        size_t ncap = have + need + w.sym.size() / 2;
        w.sym.resize(ncap > max_syms ? max_syms : ncap);
        return true;
    };
    bool decoded_ok = false;
    for (long long cand = w.scan_from; cand < scan_limit && cand + 64 < total_bits; ++cand) {
        if ((cand & 1023) == 0 && stop_asked()) break;
        Stream s;
        seek_bit(s, c->base, c->end, cand);
        if (s.cnt < 17 || s.peek(3) != 4) continue;       // BFINAL = 0, BTYPE = 2 (dynamic codes)
        s.drop(3);
        if (!read_dynamic_tables(s, dyn)) continue;
        // a header that parses: decode from here
        size_t have = 0;
        bool first = true, fail = false;
        const Ent *lt = dyn->lit, *dt = dyn->dist;
        int last = 0, kind = 2;      // the block at hand: 0 = stored (already copied), 1 / 2 = coded with lt / dt
        w.next = -1, w.finished = false, w.full = false;
        for (;;) {
            while (kind != 0) {
                uint16_t* o = w.sym.data() + have;
                const int r = decode_block_sym(s, lt, dt, w.sym.data(), o, w.sym.data() + w.sym.size());
                have = (size_t)(o - w.sym.data());
                if (r == INF_OK) break;
                if (r == INF_FULL) {
                    if (room(have, SYM_ROOM)) continue;
                    w.full = true;
                    break;
                }
                fail = true;
                break;
            }
            if (fail) break;
            if (w.full) {
                if (first) fail = true;      // (full inside an unproven first block: not worth publishing)
                break;
            }
            // The start is published -- once, for good -- when the block behind the first one has a header that parses
            // too (or the stream ends / another worker's start is reached there): a false start may survive one block
            // of garbage, hardly two.
            auto publish = [&]() {
                if (first) {
                    first = false;
                    w.start_bit.store(cand, std::memory_order_release);
                }
            };
            if (last) {
                publish();
                w.finished = true;
                break;
            }
            // block boundary
            if (stop_asked()) {
                fail = true;
                break;
            }
            s.refill();
            const int j = spec_handover(c, k, bit_pos(s, c->base), w.have_off.load(std::memory_order_acquire) != 0);
            if (j >= 0) {
                publish();
                w.next = j;
                break;
            }
            if (s.cnt < 3) {
                fail = true;
                break;
            }
            last = s.peek(1);
            s.drop(1);
            kind = s.peek(2);
            s.drop(2);
            if (kind == 0) {
                s.drop(s.cnt & 7);
                if (s.cnt < 32) s.refill();
                if (s.cnt < 32) {
                    fail = true;
                    break;
                }
                const uint32_t len = s.peek(16);
                s.drop(16);
                const uint32_t nlen = s.peek(16);
                s.drop(16);
                s.p -= s.cnt >> 3;
                s.buf = 0, s.cnt = 0;
                if ((len ^ 0xffffu) != nlen || (size_t)(s.end - s.p) < len) {
                    fail = true;
                    break;
                }
                size_t take = len;
                if (!room(have, len)) {
                    room(have, max_syms - have);
                    take = w.sym.size() - have;
                    w.full = true;
                }
                uint16_t* o = w.sym.data() + have;
                for (size_t i = 0; i < take; ++i) o[i] = s.p[i];
                have += take;
                s.p += len;
                if (w.full) {
                    publish();
                    break;
                }
            } else if (kind == 1) {
                lt = fixed.t.lit, dt = fixed.t.dist;
            } else if (kind == 2) {
                if (!read_dynamic_tables(s, dyn)) {
                    fail = true;
                    break;
                }
                lt = dyn->lit, dt = dyn->dist;
            } else {
                fail = true;
                break;
            }
            publish();
        }
        if (fail && first) continue;      // not a block start after all (or asked to stop): keep looking / leave
        w.len = have;
        w.end_bit = bit_pos(s, c->base);
        decoded_ok = !fail;               // failed behind a published start: whoever arrives here goes on alone
        break;
    }
    delete dyn;
    w.ok = decoded_ok;
    if (!decoded_ok) {
        w.resolved.store(1, std::memory_order_release);
        return;
    }
    // wait for the decoder in front to arrive (or for the end of the whole job)
    while (!w.have_off.load(std::memory_order_acquire)) {
        if (c->all_done.load(std::memory_order_acquire)) return;
        std::this_thread::yield();
    }
    const size_t off = w.off;
    size_t n = w.len;
    w.full = off + n > c->out_cap;         // the stream wants more room than there is: what fits is taken, as the plain decoder does
    if (w.full) n = c->out_cap - off;
    if (w.next >= 0 && !w.full) {          // pass the place on at once: the next worker converts its literals meanwhile
        SpecWorker& nx = c->w[w.next];
        nx.pred = k;
        nx.off = off + n;
        nx.have_off.store(1, std::memory_order_release);
    }
    uint8_t* dst = c->out + off;
    const uint16_t* sy = w.sym.data();
    std::vector<uint32_t> marks;
    for (size_t i = 0; i < n; ++i) {
        const uint16_t v = sy[i];
        if (v < 256) dst[i] = (uint8_t)v;
        else marks.push_back((uint32_t)i);
    }
    if (!marks.empty()) {
        // window bytes = the finished output in front: the decoder that arrived here must have made its own bytes final
        while (w.pred > 0 && !c->w[w.pred].resolved.load(std::memory_order_acquire)) {
            if (c->all_done.load(std::memory_order_acquire)) return;
            std::this_thread::yield();
        }
        for (uint32_t i : marks) {
            const long long src = (long long)off - 32768 + (sy[i] & 0x7fff);
            if (src < 0) {                 // distance too far back
                w.ok = false;
                break;
            }
            dst[i] = c->out[src];
        }
    }
    if (w.ok) {
        w.bytes = n;
        w.adler = adler32(dst, n);
    }
    w.resolved.store(1, std::memory_order_release);
}

// The whole stream with `threads` decoders.  Returns what inflate_raw returns; *adler = check value of the output when
// *adler_known.  `s` is left behind the last block that was decoded.
int inflate_mt(Stream& s, uint8_t* out, size_t out_cap, size_t* out_len, bool* finished, std::atomic<size_t>* progress,
               int threads, uint32_t* adler, bool* adler_known) {
    *adler_known = false;
    SpecCtl c;
    c.base = s.p, c.end = s.end, c.out = out, c.out_cap = out_cap, c.n = threads;
    std::vector<SpecWorker> workers(threads);
    c.w = workers.data();
    const size_t in_len = (size_t)(s.end - s.p);
    std::vector<std::thread> pool;
    for (int k = 1; k < threads; ++k) workers[k].scan_from = (long long)(in_len / threads * k) * 8;
    for (int k = 1; k < threads; ++k) {
        try {
            pool.emplace_back(spec_run, &c, k);
        } catch (...) {      // no more threads to be had: the workers that did start (and the plain decoder) do the job
            for (int j = k; j < threads; ++j) workers[j].cancel.store(1, std::memory_order_release);
            break;
        }
    }
    int handed = -1;
    size_t len0 = 0;
    int rc = inflate_raw(s, out, 0, out_cap, &len0, finished, progress, &c, &handed);
    size_t off = len0;
    if (rc == INF_HANDED) {
        uint32_t ad = adler32(out, len0);
        bool ad_ok = true;
        int cur = handed;
        workers[cur].pred = 0;
        workers[cur].off = off;
        workers[cur].have_off.store(1, std::memory_order_release);
        for (;;) {
            SpecWorker& w = workers[cur];
            while (!w.resolved.load(std::memory_order_acquire)) std::this_thread::yield();
            if (!w.ok) {
                // the worker failed behind its start: the plain decoder goes on from there, with the real window
                c.all_done.store(1, std::memory_order_release);
                seek_bit(s, c.base, c.end, w.start_bit.load(std::memory_order_acquire));
                size_t len1 = 0;
                rc = inflate_raw(s, out, off, out_cap, &len1, finished, progress);
                off = len1;
                ad_ok = false;
                break;
            }
            off += w.bytes;
            ad = adler32_combine(ad, w.adler, w.bytes);
            if (progress) progress->store(off, std::memory_order_release);
            if (w.full) {
                rc = INF_FULL;
                ad_ok = false;
                break;
            }
            if (w.finished) {
                rc = INF_OK;
                *finished = true;
                seek_bit(s, c.base, c.end, w.end_bit);
                break;
            }
            cur = w.next;     // (a worker that is ok, not full and not finished has handed over)
        }
        if (ad_ok && rc == INF_OK) *adler = ad, *adler_known = true;
    }
    c.all_done.store(1, std::memory_order_release);
    for (auto& t : pool) t.join();
    *out_len = off;
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
                               std::atomic<size_t>* progress, int threads) {
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
    // several decoders only where the stream is long enough to pay for their start (128 KB of compressed data each)
    if (threads > 8) threads = 8;
    while (threads > 1 && in_len / threads < (size_t(128) << 10)) --threads;
    uint32_t adler_mt = 1;
    bool adler_known = false;
    const int rc = threads > 1 ? inflate_mt(s, out, out_cap, out_len, &finished, progress, threads, &adler_mt, &adler_known)
                               : inflate_raw(s, out, 0, out_cap, out_len, &finished, progress);
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
        if ((adler_known ? adler_mt : adler32(out, *out_len)) != want) {
            llfe_set_error("llfe_inflate_zlib: incorrect data check");
            return LLFE_E_INVALID;
        }
    }
    if (progress) progress->store(*out_len, std::memory_order_release);
    return LLFE_OK;
}

extern "C" int llfe_inflate_zlib(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len) {
    return llfe_inflate_zlib_progress(in, in_len, out, out_cap, out_len, nullptr, 1);
}

extern "C" int llfe_inflate_zlib_mt(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len, int threads) {
    return llfe_inflate_zlib_progress(in, in_len, out, out_cap, out_len, nullptr, threads < 1 ? 1 : threads);
}
