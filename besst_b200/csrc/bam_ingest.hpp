// bam_ingest.hpp -- the host side of the device BAM ingest (besst_bam_ingest, include/besst_b200.h): BGZF block table,
// window loop, BAM header, the VERIFIED record-boundary chain and the bookkeeping around the kernels.
//
// Written against a Backend so the same loop drives the CUDA kernels (besst_bamdev.cu: double-buffered windows, copies
// and kernels overlapped on two streams) and a lane-serial host rendering of the same kernels (bgzf_hostcheck.cpp) that
// the CPU test suite checks against zlib and the pure-Python BAM reader: only the launch plumbing is GPU-only.
//
// One window = the BGZF blocks found in up to `window_bytes` of the file (bounded also by `max_inflated`):
//   upload       compressed bytes + block table -> device                                   (async, copy stream)
//   inflate      one warp per block: raw deflate -> its slot of the window's inflated buffer, CRC-32 checked
//   scan         one warp per block: seed = first believable record start (the block start for htslib-written files,
//                which never split a record across blocks), then hop over the block_size fields: record offsets, count,
//                landing point
//   verify       HOST, O(blocks): walk the blocks in order -- every block's seed must be exactly where the chain of the
//                preceding blocks landed; a block with a wrong seed is re-hopped from the right offset (`rescan`), blocks
//                lying inside one long record are skipped.  The accepted chain IS the sequential chain.
//   decode       one warp per block, one record per lane: fixed core + CIGAR lengths -> the record columns at the
//                block's record base (coalesced)
//   carry        the bytes of a record cut by the window's end move in front of the next window's data
#pragma once

#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

namespace bamingest {

struct BlockEntry {      // one BGZF block of a window (device-visible, 5 words)
    uint32_t cin;        // offset of its deflate stream in the window's compressed bytes
    uint32_t clen;
    uint32_t usize;
    uint32_t crc;
    uint32_t out;        // offset of its inflated bytes in the window's inflated buffer
};

struct ScanEntry {       // result of the scan kernel for one block
    uint32_t seed;       // where its hop started
    uint32_t land;       // where it ended: first record start at or past the block's end, or the cut record (PARTIAL)
    uint32_t count;      // records starting in the block (all complete inside the window)
    uint32_t flags;      // bgzf::SCAN_*
};

struct DecodeEntry {     // per block, made by the host verification
    uint32_t count;      // records to decode (0: the block lies inside a long record)
    uint32_t pad;
    int64_t base;        // ordinal of its first record in the file
};

struct Options {
    int64_t window_bytes = 512ll << 20;    // compressed bytes per window (~25 k BGZF blocks: 3.5 waves of one warp per block)
    int64_t max_inflated = 3072ll << 20;   // inflated bytes per window (offsets are 32-bit)
    int64_t carry_max = 4ll << 20;         // longest record tail a window may hand to the next one
    int64_t head_records = 1000;
    bool check_crc = true;
    // one PART of the file (multi-GPU ingest: rank r of n).  With D = the file offset of the BGZF block in which the header
    // ends, part p owns the blocks that START inside [D + (size - D) p / n, D + (size - D) (p + 1) / n) and the records that
    // start in those blocks; its last record may end in the blocks behind the range: up to `tail_bytes` of them are inflated
    // with the last window, not decoded.
    int32_t part = 0, n_parts = 1;
    int64_t tail_bytes = 1ll << 20;
    // BGZF virtual offset (block file offset << 16 | offset inside the inflated block) of the part's first record when the
    // caller knows it (the previous part's landing); -1: the end of the header if it lies in the part's first block, else
    // the first believable record start is trusted -- the caller compares first_voffset with the previous part's landing
    int64_t start_voffset = -1;
};

struct Stats {
    int64_t compressed_bytes = 0, uncompressed_bytes = 0, blocks = 0, records = 0, windows = 0, rescans = 0;
};

struct Window {
    std::vector<BlockEntry> blocks;   // non-empty blocks only
    std::vector<int64_t> fpos;        // file offset of each of them
    int64_t n_owned = 0;              // blocks starting before the range's end (the others are the tail)
    int64_t file_off = 0;             // first byte of the window in the file
    int64_t consumed = 0;             // bytes of whole blocks
    int64_t inflated = 0;             // sum of usize
    bool last = false;                // reaches the end of the file
};

static inline uint32_t rd16(const unsigned char* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }
static inline uint32_t rd32(const unsigned char* p) { return rd16(p) | rd16(p + 2) << 16; }

// BGZF block headers (SAM spec 4.1) of the bytes [0, have) read at file offset file_off.  -> false: corrupt
static inline bool scan_block_headers(const unsigned char* f, int64_t have, int64_t file_off, int64_t file_size, int64_t range_hi,
                                      int64_t data_base, const Options& opt, Window* w, int64_t* n_all_blocks, std::string* why) {
    w->blocks.clear();
    w->fpos.clear();
    w->n_owned = 0;
    w->file_off = file_off;
    w->inflated = 0;
    int64_t o = 0;
    while (o < have) {
        if (have - o < 18) break;   // header incomplete: next window
        if (f[o] != 0x1f || f[o + 1] != 0x8b || f[o + 2] != 8 || !(f[o + 3] & 4)) { *why = "not a BGZF file (bad gzip member header)"; return false; }
        const uint32_t xlen = rd16(f + o + 10);
        if (have - o < 12 + (int64_t)xlen) break;
        uint32_t bsize = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) {
            const unsigned char* sf = f + o + 12 + x;
            const uint32_t slen = rd16(sf + 2);
            if (sf[0] == 'B' && sf[1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = rd16(sf + 4) + 1;
            x += 4 + slen;
        }
        if (bsize == 0 || bsize < 12 + xlen + 8) { *why = "corrupt BGZF block header"; return false; }
        if (o + bsize > have) break;   // block incomplete: next window
        BlockEntry b;
        b.cin = (uint32_t)(o + 12 + xlen);
        b.clen = bsize - 12 - xlen - 8;
        b.usize = rd32(f + o + bsize - 4);
        b.crc = rd32(f + o + bsize - 8);
        if (b.usize > 65536u) { *why = "corrupt BGZF block (ISIZE > 64 KB)"; return false; }
        if (b.usize) {
            if (w->inflated + b.usize > opt.max_inflated && !w->blocks.empty()) break;
            b.out = (uint32_t)(data_base + w->inflated);
            w->inflated += b.usize;
            w->blocks.push_back(b);
            w->fpos.push_back(file_off + o);
            if (file_off + o < range_hi) w->n_owned = (int64_t)w->blocks.size();
        }
        ++*n_all_blocks;
        o += bsize;
    }
    w->consumed = o;
    w->last = file_off + o >= range_hi || file_off + o >= file_size;
    if (o == 0 && have > 0 && file_off + have >= file_size) {
        *why = "truncated BGZF block at the end of the file";
        return false;
    }
    return true;   // consumed == 0 with more file behind: the first block is larger than what was read (the caller reads more)
}

// BAM header (SAM spec 4.2) from the first `have` inflated bytes.  -> 1 parsed (*end = first record), 0 need more bytes,
// -1 corrupt
static inline int parse_header(const unsigned char* d, int64_t have, std::vector<std::string>* names, std::vector<int64_t>* lens,
                               int64_t* end, std::string* why) {
    if (have < 12) return 0;
    if (memcmp(d, "BAM\1", 4) != 0) { *why = "not a BAM file (bad magic)"; return -1; }
    const int64_t l_text = (int32_t)rd32(d + 4);
    if (l_text < 0) { *why = "corrupt BAM header"; return -1; }
    int64_t o = 8 + l_text;
    if (have < o + 4) return 0;
    const int64_t n_ref = (int32_t)rd32(d + o);
    if (n_ref < 0) { *why = "corrupt BAM header"; return -1; }
    o += 4;
    names->clear();
    lens->clear();
    for (int64_t r = 0; r < n_ref; ++r) {
        if (have < o + 4) return 0;
        const int64_t l_name = (int32_t)rd32(d + o);
        if (l_name < 1) { *why = "corrupt BAM reference name"; return -1; }
        if (have < o + 4 + l_name + 4) return 0;
        names->emplace_back(reinterpret_cast<const char*>(d + o + 4), (size_t)l_name - 1);
        lens->push_back((int32_t)rd32(d + o + 4 + l_name));
        o += 4 + l_name + 4;
    }
    *end = o;
    return 1;
}

struct Result {
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lengths;
    int64_t n_records = 0, n_head = 0;
    int64_t first_voffset = -1;     // virtual offset of the part's first record (-1: the part holds none)
    int64_t landing_voffset = -1;   // virtual offset of the first record BEHIND the part: the next part's first record
                                    // (-2: the chain that started at a GUESSED first record broke -- the guess was wrong,
                                    // the part holds no records and has to be read again from the previous part's landing)
    Stats stats;
};

constexpr int RC_OK = 0, RC_ERROR = -1, RC_WINDOW_TOO_SMALL = -2;

// Backend concept:
//   int64_t file_size();
//   bool load(int buf, int64_t file_off, int64_t want, const unsigned char** bytes, int64_t* have)   read into staging[buf]
//   bool upload(int buf, const Window&)                      staging + block table -> device (async)
//   bool inflate(int buf, const Window&, bool check_crc)     launch; errors surface in scan_results
//   bool read_inflated(int buf, int64_t off, int64_t n, unsigned char* dst)   (sync) for the header
//   bool scan(int buf, const Window&, int64_t cur, int64_t wend, int32_t n_ref)   launch scan (cur < 0: every block guesses its seed), queue the result copy
//   bool scan_results(int buf, const Window&, ScanEntry** entries, std::string* why)   (sync) + inflate/CRC verdicts
//   bool rescan(int buf, const Window&, int64_t k, int64_t start, int64_t wend, ScanEntry* e)   (sync) re-hop one block
//   bool decode(int buf, const Window&, const std::vector<DecodeEntry>&, int64_t n_before, int64_t n_after, int64_t est_total)
//   bool carry(int from_buf, int64_t src_off, int64_t n, int to_buf, int64_t dst_off)
//   bool inflate_verdict(int buf, std::string* why)           (sync) did the last inflate of this buffer succeed (a part's header window)
//   bool finish(std::string* why, bool* bad_records)          (sync) decode verdicts; *bad_records: a record's name / CIGAR overran it
//   std::string error()
// first BGZF block header at or behind `lo`: magic + BC subfield, and the two blocks chained behind it must look the same
template <class Backend>
int64_t find_block_start(Backend& B, int64_t lo, int64_t fsize, std::string* why) {
    if (lo <= 0) return 0;
    if (lo >= fsize) return fsize;
    const unsigned char* f = nullptr;
    int64_t have = 0;
    if (!B.load(0, lo, 4 * 65536 + 64, &f, &have)) { *why = B.error(); return -1; }
    auto block_at = [&](int64_t o) -> int64_t {   // total size of a believable block at o, 0 if none, -1 if cut by the buffer
        if (have - o < 18) return -1;
        if (f[o] != 0x1f || f[o + 1] != 0x8b || f[o + 2] != 8 || !(f[o + 3] & 4)) return 0;
        const uint32_t xlen = rd16(f + o + 10);
        if (have - o < 12 + (int64_t)xlen) return -1;
        uint32_t bsize = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) {
            const unsigned char* sf = f + o + 12 + x;
            const uint32_t slen = rd16(sf + 2);
            if (sf[0] == 'B' && sf[1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = rd16(sf + 4) + 1;
            x += 4 + slen;
        }
        return (bsize == 0 || bsize < 12 + xlen + 8) ? 0 : (int64_t)bsize;
    };
    for (int64_t o = 0; o + 18 <= have; ++o) {
        int64_t p = o;
        int ok = 0;
        for (; ok < 3; ++ok) {
            if (lo + p == fsize) { ok = 3; break; }     // chained exactly to the end of the file
            const int64_t bs = block_at(p);
            if (bs == 0) break;
            if (bs < 0) { ok = ok ? 3 : 0; break; }     // ran out of buffer after at least one whole block
            p += bs;
        }
        if (ok == 3) return lo + o;
    }
    if (lo + have >= fsize) return fsize;   // nothing but a block's tail up to the end of the file
    *why = "no BGZF block header found behind offset " + std::to_string(lo);
    return -1;
}

template <class Backend>
int run(Backend& B, const Options& opt, Result* res, std::string* why) {
    const int64_t fsize = B.file_size();
    const int64_t BASE = opt.carry_max;   // the window's data starts here in the inflated buffer; a carried tail ends here
    const bool whole = opt.n_parts <= 1;
    Stats& st = res->stats;
    st = Stats();
    res->n_records = 0;
    res->n_head = 0;
    res->first_voffset = res->landing_voffset = -1;
    Window win[2];
    int64_t foff = 0;
    int64_t hi = fsize;         // the part owns the blocks starting before hi ...
    int64_t read_end = fsize;   // ... and may read up to here (hi + tail)

    // next window: reads [foff, min(foff + window, rd_end)); blocks starting at or behind own_end are its tail
    auto prepare = [&](int buf, int64_t window, int64_t rd_end, int64_t own_end) -> int {   // 1 ready, 0 nothing left, -1 error
        for (;;) {
            if (foff >= own_end || foff >= fsize) return 0;
            const unsigned char* bytes = nullptr;
            int64_t have = 0;
            // a window that reaches the end of the range takes the tail with it
            int64_t want = (foff + window >= own_end || window > rd_end - foff) ? rd_end - foff : window;
            if (!B.load(buf, foff, want, &bytes, &have)) { *why = B.error(); return -1; }
            if (!scan_block_headers(bytes, have, foff, fsize, own_end, BASE, opt, &win[buf], &st.blocks, why)) return -1;
            if (win[buf].consumed == 0) {   // a window smaller than one BGZF block (<= 64 KB): read a whole block's worth
                want = fsize - foff < (1ll << 17) ? fsize - foff : (1ll << 17);
                if (want <= have) { *why = "corrupt BGZF block (longer than 64 KB)"; return -1; }
                if (!B.load(buf, foff, want, &bytes, &have)) { *why = B.error(); return -1; }
                if (!scan_block_headers(bytes, have, foff, fsize, own_end, BASE, opt, &win[buf], &st.blocks, why)) return -1;
                if (win[buf].consumed == 0) { *why = "corrupt BGZF block (longer than 64 KB)"; return -1; }
            }
            foff += win[buf].consumed;
            if (win[buf].n_owned == 0) {   // nothing but empty blocks (the EOF marker) or tail blocks
                if (win[buf].last) return 0;
                continue;
            }
            if (!B.upload(buf, win[buf])) { *why = B.error(); return -1; }
            return 1;
        }
    };
    auto voffset = [&](const Window& W, int64_t pos, int64_t wend) -> int64_t {   // inflated-buffer position -> BGZF virtual offset
        if (pos >= wend) return (W.file_off + W.consumed) << 16;
        size_t a = 0, b = W.blocks.size();   // last block with out <= pos
        while (b - a > 1) {
            const size_t m = (a + b) / 2;
            if ((int64_t)W.blocks[m].out <= pos) a = m; else b = m;
        }
        return W.fpos[a] << 16 | (pos - (int64_t)W.blocks[a].out);
    };
    auto from_voffset = [&](const Window& W, int64_t v, int64_t wend) -> int64_t {   // -1: not in this window
        const int64_t cpos = v >> 16, upos = v & 0xffff;
        for (size_t k = 0; k < W.blocks.size() && W.fpos[k] <= cpos; ++k)
            if (W.fpos[k] == cpos) return (int64_t)W.blocks[k].out + upos;
        return (cpos == W.file_off + W.consumed && upos == 0) ? wend : -1;
    };

    if (fsize == 0) { *why = "empty file"; return RC_ERROR; }
    // ---- header: at the start of the file.  The whole-file run reads it from its first data window; a part reads a small
    //      window of its own first (every part needs n_ref and the place where the records begin) ------------------------------
    int cur_buf = 0;
    int64_t header_end_v = -1;
    for (int64_t hw = whole ? opt.window_bytes : (opt.window_bytes < (8ll << 20) ? opt.window_bytes : (8ll << 20));; hw *= 4) {
        foff = 0;
        st.blocks = 0;
        const int hv = prepare(cur_buf, hw, fsize, fsize);
        if (hv < 0) return RC_ERROR;
        if (hv == 0) { *why = "no BGZF data blocks (not a BAM file)"; return RC_ERROR; }
        if (!B.inflate(cur_buf, win[cur_buf], opt.check_crc)) { *why = B.error(); return RC_ERROR; }
        std::vector<unsigned char> hb;
        int64_t got = 0, hend = 0;
        const int64_t avail = win[cur_buf].inflated;
        int rc = 0;
        for (int64_t want = 1 << 20;; want *= 4) {
            const int64_t n = want < avail ? want : avail;
            hb.resize((size_t)n);
            if (!B.read_inflated(cur_buf, BASE + got, n - got, hb.data() + got)) { *why = B.error(); return RC_ERROR; }
            got = n;
            rc = parse_header(hb.data(), got, &res->ref_names, &res->ref_lengths, &hend, why);
            if (rc != 0 || got == avail) break;
        }
        if (rc < 0) return RC_ERROR;
        if (rc == 1) {
            header_end_v = voffset(win[cur_buf], BASE + hend, BASE + avail);
            break;
        }
        if (win[cur_buf].last) { *why = "truncated BAM header"; return RC_ERROR; }
        if (whole) return RC_WINDOW_TOO_SMALL;   // the caller retries with a larger first window
    }
    const int32_t n_ref = (int32_t)res->ref_names.size();
    int64_t cur;    // first unconsumed byte of the current window's inflated buffer; -1: not known (trust the first seed)
    int64_t wend;
    if (whole) {
        wend = BASE + win[cur_buf].inflated;
        cur = from_voffset(win[cur_buf], header_end_v, wend);
    } else {
        {   // the header window's inflate may have failed: surface it before moving on (its records are not used)
            std::string herr;
            if (!B.inflate_verdict(cur_buf, &herr)) { *why = herr; return RC_ERROR; }
        }
        const int64_t D = header_end_v >> 16;   // the block in which the records begin
        const int64_t lo = D + (fsize - D) / opt.n_parts * opt.part + ((fsize - D) % opt.n_parts) * opt.part / opt.n_parts;
        hi = opt.part + 1 >= opt.n_parts ? fsize
                                         : D + (fsize - D) / opt.n_parts * (opt.part + 1) + ((fsize - D) % opt.n_parts) * (opt.part + 1) / opt.n_parts;
        read_end = hi >= fsize ? fsize : (hi + opt.tail_bytes < fsize ? hi + opt.tail_bytes : fsize);
        const int64_t first = lo <= D ? D : find_block_start(B, lo, fsize, why);
        if (first < 0) return RC_ERROR;
        foff = first;
        st.blocks = 0;
        const int hv = prepare(cur_buf, opt.window_bytes, read_end, hi);
        if (hv < 0) return RC_ERROR;
        if (hv == 0) {   // no block starts inside the range: an empty part
            bool bad = false;
            if (!B.finish(why, &bad)) return RC_ERROR;
            return RC_OK;
        }
        if (!B.inflate(cur_buf, win[cur_buf], opt.check_crc)) { *why = B.error(); return RC_ERROR; }
        wend = BASE + win[cur_buf].inflated;
        cur = -1;
        if (opt.start_voffset >= 0) {   // the caller knows where the part's first record starts
            cur = from_voffset(win[cur_buf], opt.start_voffset, wend);
            if (cur < 0) { *why = "start_voffset does not point into the part's first window"; return RC_ERROR; }
        } else if (first == D) {
            cur = from_voffset(win[cur_buf], header_end_v, wend);
        }
    }
    if (!B.scan(cur_buf, win[cur_buf], cur, wend, n_ref)) { *why = B.error(); return RC_ERROR; }
    const int64_t range_begin = win[cur_buf].file_off;
    // a chain that starts at a guessed record start and then breaks says the guess was wrong, not that the file is corrupt
    const bool guessed = cur < 0;
    auto chain_error = [&](const char* msg) -> int {
        if (!guessed) { *why = msg; return RC_ERROR; }
        bool bad = false;
        std::string ignored;
        B.finish(&ignored, &bad);
        res->n_records = 0;
        res->landing_voffset = -2;
        st.records = 0;
        return RC_OK;
    };

    std::vector<DecodeEntry> dec;
    int64_t est_total = 0;
    for (;;) {
        const Window& W = win[cur_buf];
        const int nxt = cur_buf ^ 1;
        const int more = W.last ? 0 : prepare(nxt, opt.window_bytes, read_end, hi);   // overlaps this window's kernels
        if (more < 0) return RC_ERROR;
        ScanEntry* se = nullptr;
        if (!B.scan_results(cur_buf, W, &se, why)) return RC_ERROR;
        // ---- verify the chain -------------------------------------------------------------------------------------------
        const int64_t nb = (int64_t)W.blocks.size();
        const int64_t own_end = W.n_owned < nb ? (int64_t)W.blocks[(size_t)W.n_owned].out : wend;   // the tail starts here
        dec.assign((size_t)nb, DecodeEntry{0, 0, 0});
        int64_t expected = cur, total = res->n_records, carry_start = -1;
        for (int64_t k = 0; k < W.n_owned; ++k) {
            const int64_t bend = (int64_t)W.blocks[(size_t)k].out + W.blocks[(size_t)k].usize;
            dec[(size_t)k].base = total;
            ScanEntry e = se[k];
            if (expected < 0) {   // first record of a part that starts behind the header: the first seed is trusted
                if (e.seed == 0xffffffffu) continue;   // the block lies inside a record of the previous part
                expected = e.seed;
            }
            if (expected >= bend) continue;   // inside a record that started earlier
            if ((int64_t)e.seed != expected) {
                if (!B.rescan(cur_buf, W, k, expected, wend, &e)) { *why = B.error(); return RC_ERROR; }
                ++st.rescans;
            }
            if (e.flags & 2u) return chain_error("corrupt BAM record (block_size < 32)");
            if (e.flags & 4u) return chain_error("corrupt BAM data (more record starts in a BGZF block than fit)");
            if (res->first_voffset < 0 && (e.count > 0 || (e.flags & 1u))) res->first_voffset = voffset(W, expected, wend);
            dec[(size_t)k].count = e.count;
            total += e.count;
            if (e.flags & 1u) { carry_start = e.land; break; }
            expected = e.land;
        }
        if (carry_start < 0) carry_start = (expected >= 0 && expected < wend) ? expected : wend;
        if (res->n_records == 0 && est_total == 0) {   // size the columns from the first window's record density
            double frac = (double)(W.file_off + W.consumed - range_begin) / (double)(hi > range_begin ? hi - range_begin : 1);
            if (!(frac > 0) || frac > 1) frac = 1;
            est_total = (int64_t)((double)(total + 1) / frac * 1.05) + 4096;
        }
        if (!B.decode(cur_buf, W, dec, res->n_records, total, est_total)) { *why = B.error(); return RC_ERROR; }
        res->n_records = total;
        st.uncompressed_bytes += W.inflated;
        st.compressed_bytes += W.consumed;
        ++st.windows;
        if (more == 0) {
            if (hi >= fsize) {   // the whole file, or its last part: nothing may be left over
                if (wend - carry_start != 0) return chain_error("truncated BAM file (partial record at the end)");
                res->landing_voffset = fsize << 16;
            } else {
                // the part's last record ends in the tail blocks: where the hop landed is the next part's first record
                if (expected >= 0 && carry_start < own_end) return chain_error("BAM record longer than the ingest's range tail");
                res->landing_voffset = expected < 0 ? -1 : voffset(W, carry_start, wend);
            }
            break;
        }
        const int64_t carry_len = wend - carry_start;
        if (carry_len > opt.carry_max) return chain_error("BAM record longer than the ingest's carry buffer");
        if (carry_len && !B.carry(cur_buf, carry_start, carry_len, nxt, BASE - carry_len)) { *why = B.error(); return RC_ERROR; }
        cur = expected < 0 ? -1 : BASE - carry_len;
        wend = BASE + win[nxt].inflated;
        if (!B.inflate(nxt, win[nxt], opt.check_crc)) { *why = B.error(); return RC_ERROR; }
        if (!B.scan(nxt, win[nxt], cur, wend, n_ref)) { *why = B.error(); return RC_ERROR; }
        cur_buf = nxt;
    }
    {
        bool bad_records = false;
        if (!B.finish(why, &bad_records)) return RC_ERROR;
        if (bad_records) {
            res->landing_voffset = -1;
            return chain_error("corrupt BAM record (name/CIGAR longer than the record)");
        }
    }
    res->n_head = res->n_records < opt.head_records ? res->n_records : opt.head_records;
    st.records = res->n_records;
    return RC_OK;
}

}  // namespace bamingest
