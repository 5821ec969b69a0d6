// bgzf_core.cuh -- the algorithmic core of the device-side BAM ingest: raw-DEFLATE decoding
// of one BGZF block, the block's CRC-32, BAM record boundaries and the fixed-core record decode.
//
// The reference walks the BAM one pysam.AlignedRead at a time (runBESST:162, libmetrics.py:63,257,293,
// CreateGraph.py:111); pysam's htslib inflates one BGZF block after the other on one core.  Here every BGZF block
// (an independent raw-deflate stream of <= 64 KB, SAM spec 4.1) is one WARP's job:
//   * the warp's leader lane decodes Huffman symbols into a 32-entry queue in shared memory (lookup tables of
//     2^10 / 2^8 entries per warp, canonical bit-by-bit decode for the rare longer codes);
//   * the whole warp then places the queue: positions by a shuffle scan, literals in one coalesced store,
//     matches one after the other with a warp-wide copy (period-`dist` indexing for overlapping matches);
//   * the CRC-32 of the block is a 32-lane Horner scheme over coalesced words (lane l owns every 32nd word and
//     multiplies its state by x^1024 per step through four 256-entry tables), folded with x^n multiplications.
//
// Everything that is serial per block is written as host/device code against a small `Warp` policy (leader(),
// bcast(), the queue placement, the cooperative copies), so the SAME source runs lane-serially on the host:
// besst_b200/csrc/bgzf_hostcheck.cpp compiles it with g++ and tests/test_bamdev.py checks it against zlib on every
// block of real and synthetic BAM files without a GPU.  Only the four warp primitives differ on the device
// (besst_bamdev.cu).
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define BGZF_HD __host__ __device__ __forceinline__
#else
#define BGZF_HD inline
#endif

namespace bgzf {

#ifndef BGZF_LIT_BITS
#define BGZF_LIT_BITS 10
#endif
#ifndef BGZF_DIST_BITS
#define BGZF_DIST_BITS 8
#endif
constexpr int LIT_BITS = BGZF_LIT_BITS;     // index bits of the literal/length lookup table
constexpr int DIST_BITS = BGZF_DIST_BITS;   // index bits of the distance lookup table (>= 7: the code-length code borrows it)
constexpr int QUEUE = 32;       // capacity of the symbol queue; a placement round takes one symbol per lane of the group

// error codes of inflate_block (negative), 0 = ok
constexpr int E_BTYPE = -1, E_STORED = -2, E_CODELEN = -3, E_LITCODE = -4, E_DISTCODE = -5, E_SYMBOL = -6, E_OVERRUN = -7,
              E_RANGE = -8, E_SIZE = -9;

// per-warp working set in shared memory (3.6 KB)
struct WarpTables {
    uint16_t lit_lut[1 << LIT_BITS];    // (code length << 12) | symbol, 0 = code longer than LIT_BITS (or unused)
    uint16_t dist_lut[1 << DIST_BITS];
    uint16_t lit_sym[288];              // symbols in canonical order (by length, then value)
    uint16_t dist_sym[32];
    uint16_t lit_cnt[16];               // codes per length
    uint16_t dist_cnt[16];
    uint32_t queue[QUEUE];              // literal byte, or match length | distance << 16
    uint8_t lens[320];                  // code lengths while a dynamic header is read
};

// LSB-first bit reader over 32-bit words (the compressed bytes sit at an arbitrary byte offset of a word-aligned buffer;
// up to 8 bytes past the end of the stream may be read -- the caller pads its buffer -- and are never interpreted: the
// bit position is checked against the end of the stream once per queue)
struct BitReader {
    const uint32_t* w;
    uint32_t next;      // next word to load
    uint64_t buf;
    int nbits;
    uint64_t end_bit;   // first bit past the stream, counted from the word base
    BGZF_HD void seek(uint64_t byte_off) {
        next = (uint32_t)(byte_off >> 2);
        const int skip = (int)(byte_off & 3) * 8;
        buf = (uint64_t)(w[next++] >> skip);
        nbits = 32 - skip;
    }
    BGZF_HD void init(const uint32_t* words, uint64_t byte_off, int64_t n_bytes) {
        w = words;
        end_bit = (byte_off + (uint64_t)n_bytes) * 8;
        seek(byte_off);
    }
    BGZF_HD uint64_t bit_pos() const { return (uint64_t)next * 32 - (uint64_t)nbits; }
    BGZF_HD bool overrun() const { return bit_pos() > end_bit; }
    BGZF_HD void refill() {
        if (nbits <= 32) {
            buf |= (uint64_t)w[next++] << nbits;
            nbits += 32;
        }
    }
    BGZF_HD uint32_t peek(int n) const { return (uint32_t)buf & ((1u << n) - 1u); }
    BGZF_HD void drop(int n) {
        buf >>= n;
        nbits -= n;
    }
    BGZF_HD uint32_t bits(int n) {
        const uint32_t v = peek(n);
        drop(n);
        return v;
    }
    // byte offset (from the word base) of the next unread byte; only meaningful on a byte boundary
    BGZF_HD uint64_t byte_pos() const { return (uint64_t)next * 4 - (uint64_t)(nbits >> 3); }
};

BGZF_HD uint32_t bitrev(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) {
        r = (r << 1) | (v & 1);
        v >>= 1;
    }
    return r;
}

// canonical Huffman code from code lengths: cnt[len], sym[] in canonical order, lut for the codes of <= `bits` bits.
// -> `left`: 0 complete, > 0 incomplete, < 0 over-subscribed
BGZF_HD int build_code(const uint8_t* lens, int n, uint16_t* cnt, uint16_t* sym, uint16_t* lut, int bits) {
    for (int l = 0; l < 16; ++l) cnt[l] = 0;
    for (int i = 0; i < n; ++i) cnt[lens[i]]++;
    for (int i = 0; i < (1 << bits); ++i) lut[i] = 0;
    int left = 1;
    for (int l = 1; l < 16; ++l) {
        left <<= 1;
        left -= cnt[l];
        if (left < 0) return left;
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + cnt[l]);
    for (int i = 0; i < n; ++i)
        if (lens[i]) sym[offs[lens[i]]++] = (uint16_t)i;
    uint32_t code = 0;
    int idx = 0;
    for (int l = 1; l <= bits; ++l) {
        for (int k = 0; k < cnt[l]; ++k) {
            const uint16_t e = (uint16_t)((l << 12) | sym[idx++]);
            for (uint32_t j = bitrev(code, l); j < (1u << bits); j += 1u << l) lut[j] = e;
            ++code;
        }
        code <<= 1;
    }
    return left;
}

// one symbol: table hit, else canonical decode bit by bit (codes longer than the table index).  -> symbol, or -1
BGZF_HD int decode_symbol(BitReader& br, const uint16_t* lut, int bits, const uint16_t* cnt, const uint16_t* sym) {
    const uint32_t e = lut[br.peek(bits)];
    if (e) {
        br.drop((int)(e >> 12));
        return (int)(e & 0xfff);
    }
    uint64_t b = br.buf;
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
        code |= (int)(b & 1);
        b >>= 1;
        const int count = cnt[len];
        if (code - count < first) {
            br.drop(len);
            return sym[index + (code - first)];
        }
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// the code tables of a fixed-Huffman block (RFC 1951 3.2.6)
BGZF_HD int fixed_tables(WarpTables* T) {
    for (int i = 0; i < 144; ++i) T->lens[i] = 8;
    for (int i = 144; i < 256; ++i) T->lens[i] = 9;
    for (int i = 256; i < 280; ++i) T->lens[i] = 7;
    for (int i = 280; i < 288; ++i) T->lens[i] = 8;
    build_code(T->lens, 288, T->lit_cnt, T->lit_sym, T->lit_lut, LIT_BITS);
    for (int i = 0; i < 30; ++i) T->lens[i] = 5;
    build_code(T->lens, 30, T->dist_cnt, T->dist_sym, T->dist_lut, DIST_BITS);
    return 0;
}

// the header of a dynamic-Huffman block (RFC 1951 3.2.7) -> both code tables
BGZF_HD int dynamic_tables(BitReader& br, WarpTables* T) {
    br.refill();
    const int hlit = (int)br.bits(5) + 257, hdist = (int)br.bits(5) + 1, hclen = (int)br.bits(4) + 4;
    if (hlit > 286 || hdist > 30) return E_CODELEN;
    // order of the code-length code lengths: 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15, five bits each
    const uint64_t order_lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 | 6ull << 35 |
                              10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
    const uint64_t order_hi = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
    uint8_t* cl = T->lens + 300;   // the 19 code-length code lengths: only read by the build_code below, before lens[] fills up
    for (int i = 0; i < 19; ++i) cl[i] = 0;
    for (int i = 0; i < hclen; ++i) {
        br.refill();
        const int s = (int)((i < 12 ? order_lo >> (5 * i) : order_hi >> (5 * (i - 12))) & 31);
        cl[s] = (uint8_t)br.bits(3);
    }
    // the code-length code borrows the distance tables (rebuilt below)
    const int left = build_code(cl, 19, T->dist_cnt, T->dist_sym, T->dist_lut, 7);
    if (left != 0 && !(left > 0 && 19 - T->dist_cnt[0] == 1)) return E_CODELEN;   // zlib: incomplete only with a single code
    int i = 0;
    const int n = hlit + hdist;
    while (i < n) {
        br.refill();
        const int s = decode_symbol(br, T->dist_lut, 7, T->dist_cnt, T->dist_sym);
        if (s < 0) return E_CODELEN;
        if (s < 16) {
            T->lens[i++] = (uint8_t)s;
        } else {
            int rep, val = 0;
            if (s == 16) {
                if (i == 0) return E_CODELEN;
                val = T->lens[i - 1];
                rep = 3 + (int)br.bits(2);
            } else if (s == 17) {
                rep = 3 + (int)br.bits(3);
            } else {
                rep = 11 + (int)br.bits(7);
            }
            if (i + rep > n) return E_CODELEN;
            while (rep--) T->lens[i++] = (uint8_t)val;
        }
        if (br.overrun()) return E_OVERRUN;
    }
    if (T->lens[256] == 0) return E_CODELEN;   // no end-of-block code
    // the literal/length lengths are lens[0 .. hlit), the distance lengths follow: build the distance code first from a
    // copy-free view (build_code only reads lens), then the literal/length code
    int l2 = build_code(T->lens + hlit, hdist, T->dist_cnt, T->dist_sym, T->dist_lut, DIST_BITS);
    if (l2 < 0 || (l2 > 0 && hdist - T->dist_cnt[0] > 1)) return E_DISTCODE;   // incomplete: only a single distance code (zlib)
    int l1 = build_code(T->lens, hlit, T->lit_cnt, T->lit_sym, T->lit_lut, LIT_BITS);
    if (l1 < 0 || (l1 > 0 && hlit - T->lit_cnt[0] > 1)) return E_LITCODE;
    return 0;
}

// leader lane: decode up to `cap` (<= QUEUE) symbols into the queue.  *state: 0 more to come, 1 end of block, < 0 error
BGZF_HD int decode_batch(BitReader& br, WarpTables* T, int cap, int* state) {
    int n = 0;
    *state = 0;
    while (n < cap) {
        br.refill();
        const int s = decode_symbol(br, T->lit_lut, LIT_BITS, T->lit_cnt, T->lit_sym);
        if (s < 0) { *state = E_SYMBOL; break; }
        if (s < 256) {
            T->queue[n++] = (uint32_t)s;
            continue;
        }
        if (s == 256) { *state = 1; break; }
        const int c = s - 257;
        if (c > 28) { *state = E_SYMBOL; break; }
        int len;
        if (c < 8) len = 3 + c;
        else if (c == 28) len = 258;
        else {
            const int eb = (c >> 2) - 1;
            len = 3 + ((4 + (c & 3)) << eb) + (int)br.bits(eb);
        }
        br.refill();
        const int d = decode_symbol(br, T->dist_lut, DIST_BITS, T->dist_cnt, T->dist_sym);
        if (d < 0 || d > 29) { *state = E_SYMBOL; break; }
        int dist;
        if (d < 4) dist = 1 + d;
        else {
            const int eb = (d >> 1) - 1;
            dist = 1 + ((2 + (d & 1)) << eb) + (int)br.bits(eb);
        }
        T->queue[n++] = (uint32_t)len | (uint32_t)dist << 16;
    }
    if (br.overrun()) *state = E_OVERRUN;
    return n;
}

// One BGZF block: the raw-deflate stream of `clen` bytes at byte offset `coff` of the word-aligned buffer `cwords`
// -> `usize` bytes at out.  Warp policy W:
//   width()                          lanes of the group that works on this block (symbols per placement round)
//   leader()                         true on the lane that runs the serial part
//   bcast(v)                         the leader's v on every lane; orders the leader's shared-memory writes before it
//   place(T, n, out, pos, usize)     put the n queued symbols at out[pos ..): -> bytes written, or -1 (range error)
//   copy_in(dst, src_bytes, n)       cooperative byte copy (stored blocks)
// -> 0, or a negative E_* code (the same on every lane)
template <class W>
BGZF_HD int inflate_block(W& wp, const uint32_t* cwords, uint64_t coff, uint32_t clen, uint8_t* out, uint32_t usize, WarpTables* T) {
    BitReader br;
    br.w = cwords; br.next = 0; br.buf = 0; br.nbits = 0; br.end_bit = 0;
    if (wp.leader()) br.init(cwords, coff, clen);
    uint32_t pos = 0;
    for (;;) {
        // ---- block header: final flag, type, tables --------------------------------------------------------------
        int hdr = 0;
        uint32_t stored_len = 0;
        uint64_t stored_src = 0;
        if (wp.leader()) {
            br.refill();
            const int bfinal = (int)br.bits(1), btype = (int)br.bits(2);
            int rc = 0;
            if (btype == 0) {
                br.drop(br.nbits & 7);   // to the byte boundary
                br.refill();
                const uint32_t len = br.bits(16), nlen = br.bits(16);
                if ((len ^ nlen) != 0xffffu) rc = E_STORED;
                stored_len = len;
                stored_src = br.byte_pos();
                if (br.bit_pos() + 8ull * len > br.end_bit) rc = E_OVERRUN;
            } else if (btype == 1) {
                rc = fixed_tables(T);
            } else if (btype == 2) {
                rc = dynamic_tables(br, T);
            } else {
                rc = E_BTYPE;
            }
            hdr = rc < 0 ? rc : (bfinal | btype << 1);
        }
        hdr = wp.bcast(hdr);
        if (hdr < 0) return hdr;
        const int bfinal = hdr & 1, btype = hdr >> 1;
        if (btype == 0) {
            stored_len = (uint32_t)wp.bcast((int)stored_len);
            const uint32_t src_lo = (uint32_t)wp.bcast((int)(uint32_t)stored_src);
            const uint32_t src_hi = (uint32_t)wp.bcast((int)(uint32_t)(stored_src >> 32));
            stored_src = (uint64_t)src_hi << 32 | src_lo;
            if (pos + stored_len > usize) return E_RANGE;
            wp.copy_in(out + pos, reinterpret_cast<const uint8_t*>(cwords) + stored_src, stored_len);
            pos += stored_len;
            if (wp.leader()) br.seek(stored_src + stored_len);
        } else {
            for (;;) {
                int n = 0, st = 0;
                if (wp.leader()) n = decode_batch(br, T, wp.width(), &st);
                const int packed = wp.bcast((n & 0xff) + st * 256);
                n = packed & 0xff;
                st = packed >> 8;   // arithmetic shift: negative states survive
                if (n) {
                    const int wrote = wp.place(T, n, out, pos, usize);
                    if (wrote < 0) return E_RANGE;
                    pos += (uint32_t)wrote;
                }
                if (st < 0) return st;
                if (st == 1) break;
            }
        }
        if (bfinal) break;
    }
    return pos == usize ? 0 : E_SIZE;
}

// ---- CRC-32 (IEEE, reflected; the gzip trailer of every BGZF block) ------------------------------------------------------
// polynomial arithmetic in the reflected representation: bit 31 is x^0
constexpr uint32_t CRC_POLY = 0xedb88320u;

// a(x) * b(x) mod P
BGZF_HD uint32_t crc_mul(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}

// x^(8n) mod P: what n zero bytes do to a CRC state
BGZF_HD uint32_t crc_xpow8n(uint64_t n) {
    uint32_t p = 1u << 31;           // x^0
    uint32_t sq = 1u << 23;          // x^8
    while (n) {
        if (n & 1) p = crc_mul(sq, p);
        sq = crc_mul(sq, sq);
        n >>= 1;
    }
    return p;
}

// tables: crc_tab[0] = the byte table T[b] (state after byte b from state 0, i.e. b(x) * x^8... the classic table);
// crc_tab[1 + k][b] = (b << 8k as a state) advanced over 128 zero bytes: one Horner step of a lane is four lookups
struct CrcTables {
    uint32_t byte_tab[256];
    uint32_t z128[4][256];
};

inline void crc_make_tables(CrcTables* t) {
    for (uint32_t b = 0; b < 256; ++b) {
        uint32_t c = b;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ CRC_POLY : c >> 1;
        t->byte_tab[b] = c;
    }
    const uint32_t x1024 = crc_xpow8n(128);
    for (int k = 0; k < 4; ++k)
        for (uint32_t b = 0; b < 256; ++b) t->z128[k][b] = crc_mul(x1024, b << (8 * k));
}

BGZF_HD uint32_t crc_z128(const CrcTables* t, uint32_t s) {
    return t->z128[0][s & 0xff] ^ t->z128[1][(s >> 8) & 0xff] ^ t->z128[2][(s >> 16) & 0xff] ^ t->z128[3][s >> 24];
}

// unaligned little-endian 32-bit load from a word-aligned buffer
BGZF_HD uint32_t ld32u(const uint32_t* words, uint64_t byte_off) {
    const uint64_t i = byte_off >> 2;
    const int sh = (int)(byte_off & 3) * 8;
    const uint32_t lo = words[i];
    if (sh == 0) return lo;
    return (lo >> sh) | (words[i + 1] << (32 - sh));
}
BGZF_HD uint32_t ld8u(const uint32_t* words, uint64_t byte_off) { return (words[byte_off >> 2] >> ((byte_off & 3) * 8)) & 0xff; }
BGZF_HD uint32_t ld16u(const uint32_t* words, uint64_t byte_off) { return ld8u(words, byte_off) | ld8u(words, byte_off + 1) << 8; }

// Lane `lane` of 32: Horner state over the words lane, lane + 32, ... of the first `rounds` 128-byte rows of the message
// at byte offset `off`; the caller folds the 32 states with crc_fold_lane and finishes the tail bytes.
BGZF_HD uint32_t crc_lane_rows(const CrcTables* t, const uint32_t* words, uint64_t off, uint32_t rounds, int lane) {
    uint32_t s = 0;
    for (uint32_t r = 0; r < rounds; ++r) s = crc_z128(t, s) ^ ld32u(words, off + 128ull * r + 4ull * lane);
    return s;
}
// the lane's contribution to the state after rounds * 128 bytes: its words still have 4 zero bytes (the word itself, as a
// state) plus 4 * (31 - lane) bytes of the last row to pass
BGZF_HD uint32_t crc_fold_lane(uint32_t s, int lane) { return crc_mul(crc_xpow8n(4ull * (32 - lane)), s); }
// state after `n_bytes` (a multiple of 128) given the XOR of the folded lane states and the initial state 0xffffffff
BGZF_HD uint32_t crc_with_init(uint32_t folded_xor, uint64_t n_bytes) { return folded_xor ^ crc_mul(crc_xpow8n(n_bytes), 0xffffffffu); }
BGZF_HD uint32_t crc_tail(const CrcTables* t, uint32_t state, const uint32_t* words, uint64_t off, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) state = t->byte_tab[(state ^ ld8u(words, off + i)) & 0xff] ^ (state >> 8);
    return state;
}

// ---- BAM records ----------------------------------------------------------------------------------------------------
// A record: block_size (int32), then the 32-byte fixed core (SAM spec 4.2): refID pos l_read_name mapq bin n_cigar flag
// l_seq next_refID next_pos tlen, read name, CIGAR, ...

// Is `o` a believable record start?  Used to seed the hop through a BGZF block that does not begin at a record boundary;
// the host verifies every seed against the chain of the preceding blocks, so a wrong guess costs time, never correctness.
BGZF_HD bool record_plausible(const uint32_t* u, uint64_t o, uint64_t end, int32_t n_ref) {
    if (o + 36 > end) return false;
    const int32_t bs = (int32_t)ld32u(u, o);
    if (bs < 32 || bs > (1 << 28)) return false;
    const int32_t ref = (int32_t)ld32u(u, o + 4), pos = (int32_t)ld32u(u, o + 8);
    const int32_t mref = (int32_t)ld32u(u, o + 24), mpos = (int32_t)ld32u(u, o + 28);
    if (ref < -1 || ref >= n_ref || mref < -1 || mref >= n_ref || pos < -1 || mpos < -1) return false;
    const uint32_t l_name = ld8u(u, o + 12), n_cigar = ld16u(u, o + 16);
    const int32_t l_seq = (int32_t)ld32u(u, o + 20);
    if (l_name < 1 || l_seq < 0) return false;
    if (32ll + l_name + 4ll * n_cigar + ((int64_t)l_seq + 1) / 2 + l_seq > bs) return false;
    if (o + 36 + l_name <= end && ld8u(u, o + 36 + l_name - 1) != 0) return false;   // NUL-terminated read name
    const uint64_t nx = o + 4 + (uint64_t)bs;   // the next record must look like one, too
    if (nx + 36 <= end) {
        const int32_t bs2 = (int32_t)ld32u(u, nx), ref2 = (int32_t)ld32u(u, nx + 4), mref2 = (int32_t)ld32u(u, nx + 24);
        if (bs2 < 32 || bs2 > (1 << 28) || ref2 < -1 || ref2 >= n_ref || mref2 < -1 || mref2 >= n_ref || ld8u(u, nx + 12) < 1) return false;
    }
    return true;
}

constexpr uint32_t SCAN_PARTIAL = 1, SCAN_CORRUPT = 2, SCAN_OVERFLOW = 4;
constexpr int MAX_RECORDS_PER_BLOCK = 2048;   // a 64 KB block holds at most 65536 / 36 = 1820 record starts

// Hop over the block_size fields from `start`: the records that START before `block_end` belong to this block; their
// offsets go to offs[].  -> where the hop lands (>= block_end), or the start of the record that is not complete inside the
// window (SCAN_PARTIAL)
BGZF_HD uint64_t hop_block(const uint32_t* u, uint64_t start, uint64_t block_end, uint64_t wend, uint32_t* offs, uint32_t* count,
                           uint32_t* flags) {
    uint64_t o = start;
    uint32_t n = 0, fl = 0;
    while (o < block_end) {
        if (o + 4 > wend) { fl |= SCAN_PARTIAL; break; }
        const int32_t bs = (int32_t)ld32u(u, o);
        if (bs < 32) { fl |= SCAN_CORRUPT; break; }
        if (o + 4 + (uint64_t)bs > wend) { fl |= SCAN_PARTIAL; break; }
        if (n >= (uint32_t)MAX_RECORDS_PER_BLOCK) { fl |= SCAN_OVERFLOW; break; }
        offs[n++] = (uint32_t)o;
        o += 4 + (uint64_t)bs;
    }
    *count = n;
    *flags = fl;
    return o;
}

struct RecordFields {
    int32_t tid, pos, mtid, mpos, tlen, qlen, rlen, alen;
    uint32_t flag, mapq;
    bool ok;
};

// fixed core + CIGAR-derived lengths with pysam 0.8.4's meaning: qlen = query_alignment_length (l_seq
// minus leading / trailing soft clips; from the CIGAR when SEQ is '*'), alen = reference span
BGZF_HD RecordFields decode_record(const uint32_t* u, uint64_t o) {
    RecordFields f;
    const int64_t bs = (int32_t)ld32u(u, o);
    f.tid = (int32_t)ld32u(u, o + 4);
    f.pos = (int32_t)ld32u(u, o + 8);
    const uint32_t w12 = ld32u(u, o + 12), w16 = ld32u(u, o + 16);
    const uint32_t l_name = w12 & 0xff, n_cigar = w16 & 0xffff;
    f.mapq = (w12 >> 8) & 0xff;
    f.flag = w16 >> 16;
    const int32_t l_seq = (int32_t)ld32u(u, o + 20);
    f.mtid = (int32_t)ld32u(u, o + 24);
    f.mpos = (int32_t)ld32u(u, o + 28);
    f.tlen = (int32_t)ld32u(u, o + 32);
    f.rlen = l_seq;
    f.ok = 32 + (int64_t)l_name + 4 * (int64_t)n_cigar <= bs;
    int64_t q_start = 0, q_end = l_seq, ref_span = 0;
    if (f.ok && n_cigar) {
        const uint64_t c0 = o + 36 + l_name;
        int64_t q_from_cigar = 0;
        bool leading = true;
        int64_t trailing = 0;   // soft-clipped bases since the last operation that is neither S nor H
        for (uint32_t k = 0; k < n_cigar; ++k) {
            const uint32_t c = ld32u(u, c0 + 4ull * k), op = c & 0xf, len = c >> 4;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_span += len;
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) q_from_cigar += len;
            if (leading) {
                if (op == 4) q_start += len;
                else if (op != 5) leading = false;
            }
            if (op == 4) trailing += len;
            else if (op != 5) trailing = 0;
        }
        if (l_seq == 0) q_end = q_from_cigar;
        if (n_cigar > 1) q_end -= trailing;
    }
    f.qlen = (int32_t)(q_end - q_start);
    f.alen = (int32_t)ref_span;
    return f;
}

}  // namespace bgzf
