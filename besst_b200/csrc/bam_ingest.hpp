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
};

struct Stats {
    int64_t compressed_bytes = 0, uncompressed_bytes = 0, blocks = 0, records = 0, windows = 0, rescans = 0;
};

struct Window {
    std::vector<BlockEntry> blocks;   // non-empty blocks only
    int64_t file_off = 0;             // first byte of the window in the file
    int64_t consumed = 0;             // bytes of whole blocks
    int64_t inflated = 0;             // sum of usize
    bool last = false;                // reaches the end of the file
};

static inline uint32_t rd16(const unsigned char* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }
static inline uint32_t rd32(const unsigned char* p) { return rd16(p) | rd16(p + 2) << 16; }

// BGZF block headers (SAM spec 4.1) of the bytes [0, have) read at file offset file_off.  -> false: corrupt
static inline bool scan_block_headers(const unsigned char* f, int64_t have, int64_t file_off, int64_t file_size, int64_t data_base,
                                      const Options& opt, Window* w, int64_t* n_all_blocks, std::string* why) {
    w->blocks.clear();
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
        }
        ++*n_all_blocks;
        o += bsize;
    }
    w->consumed = o;
    w->last = file_off + o >= file_size;
    if (o == 0 && have > 0) {
        *why = file_off + have >= file_size ? "truncated BGZF block at the end of the file" : "BGZF block larger than the window";
        return false;
    }
    return true;
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
    Stats stats;
};

constexpr int RC_OK = 0, RC_ERROR = -1, RC_WINDOW_TOO_SMALL = -2;

// Backend concept:
//   int64_t file_size();
//   bool load(int buf, int64_t file_off, int64_t want, const unsigned char** bytes, int64_t* have)   read into staging[buf]
//   bool upload(int buf, const Window&)                      staging + block table -> device (async)
//   bool inflate(int buf, const Window&, bool check_crc)     launch; errors surface in scan_results
//   bool read_inflated(int buf, int64_t off, int64_t n, unsigned char* dst)   (sync) for the header
//   bool scan(int buf, const Window&, int64_t cur, int64_t wend, int32_t n_ref)   launch scan, queue the result copy
//   bool scan_results(int buf, const Window&, ScanEntry** entries, std::string* why)   (sync) + inflate/CRC verdicts
//   bool rescan(int buf, const Window&, int64_t k, int64_t start, int64_t wend, ScanEntry* e)   (sync) re-hop one block
//   bool decode(int buf, const Window&, const std::vector<DecodeEntry>&, int64_t n_before, int64_t n_after, int64_t est_total)
//   bool carry(int from_buf, int64_t src_off, int64_t n, int to_buf, int64_t dst_off)
//   bool finish(std::string* why)                            (sync) decode verdicts
//   std::string error()
template <class Backend>
int run(Backend& B, const Options& opt, Result* res, std::string* why) {
    const int64_t fsize = B.file_size();
    const int64_t BASE = opt.carry_max;   // the window's data starts here in the inflated buffer; a carried tail ends here
    Stats& st = res->stats;
    st = Stats();
    st.compressed_bytes = fsize;
    res->n_records = 0;
    res->n_head = 0;
    Window win[2];
    int64_t foff = 0;
    bool eof = fsize == 0;

    auto prepare = [&](int buf) -> int {   // 1 window ready, 0 nothing left, -1 error
        for (;;) {
            if (foff >= fsize) return 0;
            const unsigned char* bytes = nullptr;
            int64_t have = 0;
            if (!B.load(buf, foff, opt.window_bytes, &bytes, &have)) { *why = B.error(); return -1; }
            if (!scan_block_headers(bytes, have, foff, fsize, BASE, opt, &win[buf], &st.blocks, why)) return -1;
            foff += win[buf].consumed;
            if (win[buf].blocks.empty()) continue;   // nothing but empty blocks (the EOF marker)
            if (!B.upload(buf, win[buf])) { *why = B.error(); return -1; }
            return 1;
        }
    };

    if (eof) { *why = "empty file"; return RC_ERROR; }
    int cur_buf = 0;
    int have_win = prepare(cur_buf);
    if (have_win < 0) return RC_ERROR;
    if (have_win == 0) { *why = "no BGZF data blocks (not a BAM file)"; return RC_ERROR; }
    if (!B.inflate(cur_buf, win[cur_buf], opt.check_crc)) { *why = B.error(); return RC_ERROR; }

    // ---- header: in the first window --------------------------------------------------------------------------------
    int64_t cur;   // first unconsumed byte of the current window's inflated buffer
    {
        std::vector<unsigned char> hb;
        int64_t got = 0, hend = 0;
        const int64_t avail = win[cur_buf].inflated;
        for (int64_t want = 1 << 20;; want *= 4) {
            const int64_t n = want < avail ? want : avail;
            hb.resize((size_t)n);
            if (!B.read_inflated(cur_buf, BASE + got, n - got, hb.data() + got)) { *why = B.error(); return RC_ERROR; }
            got = n;
            const int rc = parse_header(hb.data(), got, &res->ref_names, &res->ref_lengths, &hend, why);
            if (rc < 0) return RC_ERROR;
            if (rc == 1) break;
            if (got == avail) {
                if (win[cur_buf].last) { *why = "truncated BAM header"; return RC_ERROR; }
                return RC_WINDOW_TOO_SMALL;   // the caller retries with a larger first window
            }
        }
        cur = BASE + hend;
    }
    const int32_t n_ref = (int32_t)res->ref_names.size();
    int64_t wend = BASE + win[cur_buf].inflated;
    if (!B.scan(cur_buf, win[cur_buf], cur, wend, n_ref)) { *why = B.error(); return RC_ERROR; }

    std::vector<DecodeEntry> dec;
    int64_t est_total = 0;
    for (;;) {
        const Window& W = win[cur_buf];
        const int nxt = cur_buf ^ 1;
        const int more = W.last ? 0 : prepare(nxt);   // host read + upload of the next window overlap this window's kernels
        if (more < 0) return RC_ERROR;
        ScanEntry* se = nullptr;
        if (!B.scan_results(cur_buf, W, &se, why)) return RC_ERROR;
        // ---- verify the chain -------------------------------------------------------------------------------------------
        const int64_t nb = (int64_t)W.blocks.size();
        dec.assign((size_t)nb, DecodeEntry{0, 0, 0});
        int64_t expected = cur, total = res->n_records, carry_start = -1;
        for (int64_t k = 0; k < nb; ++k) {
            const int64_t bend = (int64_t)W.blocks[(size_t)k].out + W.blocks[(size_t)k].usize;
            dec[(size_t)k].base = total;
            if (expected >= bend) continue;   // inside a record that started earlier
            ScanEntry e = se[k];
            if ((int64_t)e.seed != expected) {
                if (!B.rescan(cur_buf, W, k, expected, wend, &e)) { *why = B.error(); return RC_ERROR; }
                ++st.rescans;
            }
            if (e.flags & 2u) { *why = "corrupt BAM record (block_size < 32)"; return RC_ERROR; }
            if (e.flags & 4u) { *why = "corrupt BAM data (more record starts in a BGZF block than fit)"; return RC_ERROR; }
            dec[(size_t)k].count = e.count;
            total += e.count;
            if (e.flags & 1u) { carry_start = e.land; break; }
            expected = e.land;
        }
        if (carry_start < 0) carry_start = expected < wend ? expected : wend;
        const int64_t carry_len = wend - carry_start;
        if (res->n_records == 0 && est_total == 0) {   // size the columns from the first window's record density
            const double frac = (double)(W.file_off + W.consumed) / (double)fsize;
            est_total = (int64_t)((double)(total + 1) / (frac > 0 ? frac : 1.0) * 1.05) + 4096;
        }
        if (!B.decode(cur_buf, W, dec, res->n_records, total, est_total)) { *why = B.error(); return RC_ERROR; }
        res->n_records = total;
        st.uncompressed_bytes += W.inflated;
        ++st.windows;
        if (more == 0) {
            if (carry_len != 0) { *why = "truncated BAM file (partial record at the end)"; return RC_ERROR; }
            break;
        }
        if (carry_len > opt.carry_max) { *why = "BAM record longer than the ingest's carry buffer"; return RC_ERROR; }
        if (carry_len && !B.carry(cur_buf, carry_start, carry_len, nxt, BASE - carry_len)) { *why = B.error(); return RC_ERROR; }
        cur = BASE - carry_len;
        wend = BASE + win[nxt].inflated;
        if (!B.inflate(nxt, win[nxt], opt.check_crc)) { *why = B.error(); return RC_ERROR; }
        if (!B.scan(nxt, win[nxt], cur, wend, n_ref)) { *why = B.error(); return RC_ERROR; }
        cur_buf = nxt;
    }
    if (!B.finish(why)) return RC_ERROR;
    res->n_head = res->n_records < opt.head_records ? res->n_records : opt.head_records;
    st.records = res->n_records;
    return RC_OK;
}

}  // namespace bamingest
