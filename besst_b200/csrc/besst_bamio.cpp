// libbesst_bamio.so -- sorted BAM file -> struct-of-arrays record columns (include/besst_bamio.h).
//
// Replaces the per-record pysam iteration in front of the hot path (runBESST:162,
// libmetrics.py:63,257,293, CreateGraph.py:111) by one pass over the file:
//   1. the BGZF block table is read from the block headers (no inflate);
//   2. blocks are independent raw-deflate streams: a pool of host threads inflates a window of blocks
//      at a time straight into one contiguous buffer (dynamic block tickets);
//   3. record boundaries are a sequential hop over the 4-byte block_size fields; the fixed-core
//      fields and the CIGAR-derived lengths are then decoded by the same threads into the column
//      arrays of besst_records (BAM order kept: thread t writes records [r_t, r_{t+1})).
// Windows keep the memory bounded (64 MB of inflated data + the columns) for BAM files of any size.
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/besst_bamio.h"

namespace {

struct Block {
    uint64_t cdata;   // file offset of the deflate stream
    uint32_t clen;    // its length
    uint32_t usize;   // inflated size (ISIZE)
    uint32_t crc;     // CRC32 of the inflated data (gzip trailer)
};

template <typename T>
struct Column {
    T* p = nullptr;
    int64_t cap = 0;
    bool ensure(int64_t n, int64_t keep) {
        if (n <= cap) return true;
        int64_t want = cap ? cap : (1 << 16);
        while (want < n) want += want / 2 + 1024;
        void* q = nullptr;
        if (posix_memalign(&q, 64, (size_t)want * sizeof(T)) != 0) return false;
        if (p && keep > 0) memcpy(q, p, (size_t)keep * sizeof(T));
        free(p);
        p = static_cast<T*>(q);
        cap = want;
        return true;
    }
    ~Column() { free(p); }
};

inline uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const unsigned char* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t rdi32(const unsigned char* p) { int32_t v; memcpy(&v, p, 4); return v; }

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct besst_bam {
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lengths;
    Column<int32_t> tid, mtid, pos, mpos, tlen, qlen, rlen, alen;
    Column<uint16_t> flag;
    Column<uint8_t> mapq;
    Column<uint32_t> packed;          // flag | mapq << 12 | qlen << 20 (besst_records.packed)
    std::atomic<int> unpackable{0};   // some record has flag >= 4096 or qlen >= 4096: the packed column is not handed out
    int64_t n = 0, n_head = 0;
    int32_t stopped = 0;   // the window callback of besst_bam_stream asked to stop
    besst_bam_stats stats;
};

namespace {

void set_err(char* err, int32_t err_len, const std::string& msg) {
    if (err && err_len > 0) {
        strncpy(err, msg.c_str(), (size_t)err_len - 1);
        err[err_len - 1] = 0;
    }
}

// BGZF block table from the gzip member headers (SAM spec 4.1): magic 1f 8b 08 04, extra subfield
// 'B' 'C' with BSIZE = total block size - 1, ISIZE in the last four bytes
bool scan_blocks(const unsigned char* f, uint64_t size, std::vector<Block>* out, std::string* why) {
    uint64_t o = 0;
    while (o < size) {
        if (size - o < 28) { *why = "truncated BGZF block header"; return false; }
        if (f[o] != 0x1f || f[o + 1] != 0x8b || f[o + 2] != 8 || !(f[o + 3] & 4)) { *why = "not a BGZF file (bad gzip member header)"; return false; }
        const uint32_t xlen = rd16(f + o + 10);
        uint32_t bsize = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) {
            const unsigned char* sf = f + o + 12 + x;
            const uint32_t slen = rd16(sf + 2);
            if (sf[0] == 'B' && sf[1] == 'C' && slen == 2) bsize = (uint32_t)rd16(sf + 4) + 1;
            x += 4 + slen;
        }
        if (bsize == 0 || o + bsize > size || bsize < 12 + xlen + 8) { *why = "corrupt BGZF block"; return false; }
        Block b;
        b.cdata = o + 12 + xlen;
        b.clen = bsize - 12 - xlen - 8;
        b.usize = rd32(f + o + bsize - 4);
        b.crc = rd32(f + o + bsize - 8);
        out->push_back(b);
        o += bsize;
    }
    return true;
}

struct Inflater {
    z_stream zs;
    bool ok;
    Inflater() {
        memset(&zs, 0, sizeof(zs));
        ok = inflateInit2(&zs, -15) == Z_OK;
    }
    ~Inflater() { if (ok) inflateEnd(&zs); }
    // check_crc: compare the gzip trailer's CRC32 with the inflated bytes (a damaged block with an intact ISIZE
    // would otherwise decode into plausible garbage records)
    bool run(const unsigned char* src, uint32_t clen, unsigned char* dst, uint32_t usize, uint32_t crc, bool check_crc) {
        if (usize == 0) return true;
        if (inflateReset(&zs) != Z_OK) return false;
        zs.next_in = const_cast<unsigned char*>(src);
        zs.avail_in = clen;
        zs.next_out = dst;
        zs.avail_out = usize;
        const int rc = inflate(&zs, Z_FINISH);
        if (!(rc == Z_STREAM_END && zs.avail_out == 0)) return false;
        return !check_crc || (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, usize) == crc;
    }
};

// pysam 0.8.4 semantics: qlen = query_alignment_length (l_seq minus leading/trailing soft
// clips, inferred from the CIGAR when SEQ is '*'), alen = reference span of the CIGAR
inline void cigar_lengths(const unsigned char* cig, uint32_t n_cigar, int32_t l_seq, int32_t* qlen, int32_t* alen) {
    int64_t q_start = 0, q_end = l_seq, ref_span = 0;
    if (n_cigar) {
        int64_t q_from_cigar = 0;
        for (uint32_t k = 0; k < n_cigar; ++k) {
            const uint32_t c = rd32(cig + 4 * k), op = c & 0xF, len = c >> 4;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_span += len;
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) q_from_cigar += len;
        }
        if (l_seq == 0) q_end = q_from_cigar;
        for (uint32_t k = 0; k < n_cigar; ++k) {   // leading soft clips (hard clips are skipped)
            const uint32_t c = rd32(cig + 4 * k), op = c & 0xF;
            if (op == 4) q_start += c >> 4;
            else if (op != 5) break;
        }
        if (n_cigar > 1)
            for (uint32_t k = n_cigar; k-- > 0;) {   // trailing soft clips
                const uint32_t c = rd32(cig + 4 * k), op = c & 0xF;
                if (op == 4) q_end -= c >> 4;
                else if (op != 5) break;
            }
    }
    *qlen = (int32_t)(q_end - q_start);
    *alen = (int32_t)ref_span;
}

}  // namespace

extern "C" int besst_bamio_abi_version(void) { return 2; }

// window_fn == nullptr: all records accumulate in the handle's columns (besst_bam_read); else every decoded window is
// handed to window_fn and its column space reused (besst_bam_stream)
static besst_bam* read_impl(const char* path, int32_t n_threads, int64_t max_records, int64_t head_records,
                            besst_bam_window_fn window_fn, void* user, char* err, int32_t err_len) {
    const double t_start = now();
    int64_t base_total = 0;   // records delivered to window_fn so far
    if (n_threads <= 0) n_threads = (int32_t)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if (head_records < 0) head_records = 0;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { set_err(err, err_len, std::string("cannot open ") + path); return nullptr; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); set_err(err, err_len, std::string("cannot stat / empty file: ") + path); return nullptr; }
    const uint64_t fsize = (uint64_t)st.st_size;
    void* map = mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) { set_err(err, err_len, std::string("mmap failed: ") + path); return nullptr; }
    madvise(map, fsize, MADV_SEQUENTIAL);
    const unsigned char* f = static_cast<const unsigned char*>(map);

    besst_bam* B = new besst_bam();
    memset(&B->stats, 0, sizeof(B->stats));
    B->stats.threads = n_threads;
    B->stats.compressed_bytes = (int64_t)fsize;
    std::string why;
    auto fail = [&](const std::string& msg) -> besst_bam* {
        set_err(err, err_len, msg + " (" + path + ")");
        munmap(map, fsize);
        delete B;
        return nullptr;
    };

    std::vector<Block> blocks;
    if (!scan_blocks(f, fsize, &blocks, &why)) return fail(why);
    B->stats.blocks = (int64_t)blocks.size();

    uint64_t WINDOW = 64ull << 20;   // inflated bytes per window
    if (const char* e = getenv("BESST_BAMIO_WINDOW")) {   // tests: force many small windows
        const long long v = atoll(e);
        if (v > 0) WINDOW = (uint64_t)v;
    }
    // inflated bytes not consumed yet (header / partial record at the front); a raw buffer: no zero fill on growth
    struct Pend {
        unsigned char* p = nullptr;
        size_t len = 0, cap = 0;
        ~Pend() { free(p); }
        bool resize(size_t n) {
            if (n > cap) {
                size_t want = n + n / 8 + 4096;
                unsigned char* q = static_cast<unsigned char*>(realloc(p, want));
                if (!q) return false;
                p = q;
                cap = want;
            }
            len = n;
            return true;
        }
        unsigned char* data() { return p; }
        size_t size() const { return len; }
        bool empty() const { return len == 0; }
    } pend;
    std::vector<uint32_t> offs;            // record starts inside `pend` for the current window
    std::vector<Inflater> inflaters((size_t)n_threads);
    for (auto& z : inflaters) if (!z.ok) return fail("zlib inflateInit2 failed");
    bool header_done = false, stop = false;
    const char* nocrc = getenv("BESST_BAMIO_NOCRC");
    const bool check_crc = !(nocrc && nocrc[0] == '1');
    size_t next_block = 0;
    while (next_block < blocks.size() && !stop) {
        // ---- inflate one window of blocks behind the leftover bytes --------------------------------
        const size_t b0 = next_block;
        uint64_t usum = 0;
        while (next_block < blocks.size() && (usum < WINDOW || next_block == b0)) usum += blocks[next_block++].usize;
        const size_t b1 = next_block;
        const size_t left = pend.size();
        if (!pend.resize(left + usum)) return fail("out of memory for the inflate window");
        std::vector<uint64_t> dst_off(b1 - b0);
        {
            uint64_t o = left;
            for (size_t k = b0; k < b1; ++k) { dst_off[k - b0] = o; o += blocks[k].usize; }
        }
        const double t0 = now();
        std::atomic<size_t> ticket(b0);
        std::atomic<int> bad(0);
        auto inflate_worker = [&](int t) {
            for (;;) {
                const size_t k = ticket.fetch_add(1);
                if (k >= b1) break;
                if (!inflaters[(size_t)t].run(f + blocks[k].cdata, blocks[k].clen, pend.data() + dst_off[k - b0], blocks[k].usize, blocks[k].crc, check_crc)) bad.store(1);
            }
        };
        {
            std::vector<std::thread> th;
            for (int t = 1; t < n_threads; ++t) th.emplace_back(inflate_worker, t);
            inflate_worker(0);
            for (auto& x : th) x.join();
        }
        if (bad.load()) return fail("inflate failed: corrupt BGZF block (bad deflate stream or CRC32 mismatch)");
        B->stats.seconds_inflate += now() - t0;
        B->stats.uncompressed_bytes += (int64_t)usum;

        // ---- header (once) ------------------------------------------------------------------------------
        size_t cur = 0;
        if (!header_done) {
            const unsigned char* d = pend.data();
            const size_t have = pend.size();
            bool complete = false;
            do {
                if (have < 12) break;
                if (memcmp(d, "BAM\1", 4) != 0) return fail("not a BAM file (bad magic)");
                const int64_t l_text = rdi32(d + 4);
                if (l_text < 0) return fail("corrupt BAM header");
                size_t o = 8 + (size_t)l_text;
                if (have < o + 4) break;
                const int64_t n_ref = rdi32(d + o);
                if (n_ref < 0) return fail("corrupt BAM header");
                o += 4;
                std::vector<std::string> names;
                std::vector<int64_t> lens;
                names.reserve((size_t)n_ref);
                lens.reserve((size_t)n_ref);
                bool enough = true;
                for (int64_t r = 0; r < n_ref; ++r) {
                    if (have < o + 4) { enough = false; break; }
                    const int64_t l_name = rdi32(d + o);
                    if (l_name < 1) return fail("corrupt BAM reference name");
                    if (have < o + 4 + (size_t)l_name + 4) { enough = false; break; }
                    names.emplace_back(reinterpret_cast<const char*>(d + o + 4), (size_t)l_name - 1);
                    lens.push_back(rdi32(d + o + 4 + l_name));
                    o += 4 + (size_t)l_name + 4;
                }
                if (!enough) break;
                B->ref_names.swap(names);
                B->ref_lengths.swap(lens);
                cur = o;
                complete = true;
            } while (false);
            if (!complete) {
                if (next_block >= blocks.size()) return fail("truncated BAM header");
                continue;   // need more windows
            }
            header_done = true;
        }

        // ---- record boundaries of this window ------------------------------------------------------------
        // The hop over the block_size fields is a dependent chain of cache misses (~80 ns each): sequentially it
        // costs more than the whole inflate.  htslib flushes a BGZF block rather than split a record across two
        // (bgzf_flush_try), so in practice every block starts at a record boundary: the threads hop through
        // disjoint groups of blocks in parallel and the assumption is VERIFIED -- every group's chain must end
        // exactly where the next group starts -- else this window is scanned sequentially.
        const double t1 = now();
        const int64_t n0 = B->n;
        std::vector<std::vector<uint32_t>> toffs((size_t)n_threads);
        std::vector<int64_t> tbase((size_t)n_threads + 1, 0);
        bool parallel_scan = false;
        if (max_records < 0 && n_threads > 1) {
            // block starts at or after `cur`, plus the end of the window
            std::vector<uint64_t> starts;
            for (size_t k = 0; k < dst_off.size(); ++k)
                if (dst_off[k] >= cur && blocks[b0 + k].usize > 0) starts.push_back(dst_off[k]);
            const uint64_t wend = pend.size();
            if (!starts.empty() && starts[0] == cur) {
                starts.push_back(wend);
                const size_t nb = starts.size() - 1;
                std::vector<size_t> gfirst((size_t)n_threads + 1);
                for (int t = 0; t <= n_threads; ++t) gfirst[(size_t)t] = nb * (size_t)t / (size_t)n_threads;
                std::atomic<int> mismatch(0);
                auto hop_worker = [&](int t) {
                    const size_t g0 = gfirst[(size_t)t], g1 = gfirst[(size_t)t + 1];
                    if (g0 == g1) return;
                    const unsigned char* d = pend.data();
                    uint64_t o = starts[g0];
                    const uint64_t end = starts[g1];
                    std::vector<uint32_t>& out = toffs[(size_t)t];
                    out.reserve((size_t)((end - o) / 200 + 16));
                    while (o + 4 <= end) {
                        const int64_t bs = rdi32(d + o);
                        if (bs < 32 || o + 4 + (uint64_t)bs > end) break;
                        out.push_back((uint32_t)o);
                        o += 4 + (uint64_t)bs;
                    }
                    if (o != end) mismatch.store(1);
                };
                {
                    std::vector<std::thread> th;
                    for (int t = 1; t < n_threads; ++t) th.emplace_back(hop_worker, t);
                    hop_worker(0);
                    for (auto& x : th) x.join();
                }
                if (!mismatch.load()) {
                    parallel_scan = true;
                    for (int t = 0; t < n_threads; ++t) tbase[(size_t)t + 1] = tbase[(size_t)t] + (int64_t)toffs[(size_t)t].size();
                    cur = wend;
                } else {
                    for (auto& v : toffs) v.clear();
                }
            }
        }
        if (!parallel_scan) {   // sequential hop; the decode below splits the records evenly
            offs.clear();
            const unsigned char* d = pend.data();
            const size_t have = pend.size();
            while (cur + 4 <= have) {
                const int64_t bs = rdi32(d + cur);
                if (bs < 32) return fail("corrupt BAM record (block_size < 32)");
                if (cur + 4 + (size_t)bs > have) break;
                offs.push_back((uint32_t)cur);
                cur += 4 + (size_t)bs;
                if (max_records >= 0 && base_total + B->n + (int64_t)offs.size() >= max_records) { stop = true; break; }
            }
            const int64_t mm = (int64_t)offs.size();
            for (int t = 0; t <= n_threads; ++t) tbase[(size_t)t] = mm * t / n_threads;
        }
        const int64_t m = tbase[(size_t)n_threads];
        if (!B->tid.ensure(n0 + m, n0) || !B->mtid.ensure(n0 + m, n0) || !B->pos.ensure(n0 + m, n0) || !B->mpos.ensure(n0 + m, n0) ||
            !B->tlen.ensure(n0 + m, n0) || !B->qlen.ensure(n0 + m, n0) || !B->flag.ensure(n0 + m, n0) || !B->mapq.ensure(n0 + m, n0) ||
            !B->packed.ensure(n0 + m, n0))
            return fail("out of memory for the record columns");
        const int64_t head_new = std::min<int64_t>(head_records, base_total + n0 + m);
        if (!B->rlen.ensure(head_new > 0 ? head_new : 1, B->n_head) || !B->alen.ensure(head_new > 0 ? head_new : 1, B->n_head))
            return fail("out of memory for the record columns");

        // ---- decode: thread t takes the records it found (parallel scan) or an even share -----------------
        std::atomic<int> bad_rec(0);
        auto decode_worker = [&](int t) {
            const int64_t r0 = tbase[(size_t)t], r1 = tbase[(size_t)t + 1];
            const uint32_t* my = parallel_scan ? toffs[(size_t)t].data() - r0 : offs.data();
            const unsigned char* d = pend.data();
            for (int64_t r = r0; r < r1; ++r) {
                const unsigned char* p = d + my[r];
                const int64_t bs = rdi32(p);
                const uint32_t l_read_name = p[12], n_cigar = rd16(p + 16);
                const int32_t l_seq = rdi32(p + 20);
                if (32 + (int64_t)l_read_name + 4 * (int64_t)n_cigar > bs) { bad_rec.store(1); continue; }
                const int64_t g = n0 + r;
                B->tid.p[g] = rdi32(p + 4);
                B->pos.p[g] = rdi32(p + 8);
                B->mapq.p[g] = p[13];
                B->flag.p[g] = rd16(p + 18);
                B->mtid.p[g] = rdi32(p + 24);
                B->mpos.p[g] = rdi32(p + 28);
                B->tlen.p[g] = rdi32(p + 32);
                int32_t ql, al;
                cigar_lengths(p + 36 + l_read_name, n_cigar, l_seq, &ql, &al);
                B->qlen.p[g] = ql;
                {
                    const uint32_t fl = rd16(p + 18);
                    if (fl >= 4096u || ql < 0 || ql >= 4096) B->unpackable.store(1);
                    B->packed.p[g] = (fl & 0xfffu) | ((uint32_t)p[13] << 12) | ((uint32_t)ql << 20);
                }
                if (base_total + g < head_records) { B->rlen.p[base_total + g] = l_seq; B->alen.p[base_total + g] = al; }
            }
        };
        {
            std::vector<std::thread> th;
            for (int t = 1; t < n_threads; ++t) th.emplace_back(decode_worker, t);
            decode_worker(0);
            for (auto& x : th) x.join();
        }
        if (bad_rec.load()) return fail("corrupt BAM record (name/CIGAR longer than the record)");
        B->n = n0 + m;
        B->n_head = head_new;
        B->stats.seconds_decode += now() - t1;
        if (window_fn) {   // streaming: hand the window over, then reuse the column space
            besst_bam_columns w;
            besst_bam_get_columns(B, &w);
            const bool go_on = window_fn(user, B, &w, base_total) == 0;
            base_total += B->n;
            B->n = 0;
            if (!go_on) { B->stopped = 1; stop = true; }   // the consumer has seen enough: not an error, the handle and its statistics are returned
        }
        // keep the unconsumed tail (a partial record) for the next window
        if (cur < pend.size()) memmove(pend.data(), pend.data() + cur, pend.size() - cur);
        pend.resize(pend.size() - cur);
    }
    if (!header_done) return fail("truncated BAM header");
    if (!stop && !pend.empty()) return fail("truncated BAM file (partial record at the end)");
    munmap(map, fsize);
    B->stats.records = base_total + B->n;
    B->stats.seconds_total = now() - t_start;
    return B;
}

extern "C" int besst_bam_stopped(const besst_bam* b) { return b ? b->stopped : 0; }

extern "C" besst_bam* besst_bam_read(const char* path, int32_t n_threads, int64_t max_records, int64_t head_records,
                                     char* err, int32_t err_len) {
    return read_impl(path, n_threads, max_records, head_records, nullptr, nullptr, err, err_len);
}

extern "C" besst_bam* besst_bam_stream(const char* path, int32_t n_threads, int64_t max_records, int64_t head_records,
                                       besst_bam_window_fn window_fn, void* user, char* err, int32_t err_len) {
    if (!window_fn) { set_err(err, err_len, "besst_bam_stream: no callback"); return nullptr; }
    return read_impl(path, n_threads, max_records, head_records, window_fn, user, err, err_len);
}

extern "C" int64_t besst_bam_n_refs(const besst_bam* b) { return b ? (int64_t)b->ref_names.size() : -1; }
extern "C" const char* besst_bam_ref_name(const besst_bam* b, int64_t i) {
    return (b && i >= 0 && i < (int64_t)b->ref_names.size()) ? b->ref_names[(size_t)i].c_str() : nullptr;
}
extern "C" int64_t besst_bam_ref_length(const besst_bam* b, int64_t i) {
    return (b && i >= 0 && i < (int64_t)b->ref_lengths.size()) ? b->ref_lengths[(size_t)i] : -1;
}
extern "C" int besst_bam_get_columns(const besst_bam* b, besst_bam_columns* out) {
    if (!b || !out) return -1;
    out->n = b->n;
    out->tid = b->tid.p; out->mtid = b->mtid.p; out->pos = b->pos.p; out->mpos = b->mpos.p; out->tlen = b->tlen.p;
    out->qlen = b->qlen.p; out->flag = b->flag.p; out->mapq = b->mapq.p; out->rlen = b->rlen.p; out->alen = b->alen.p;
    out->packed = b->unpackable.load() ? nullptr : b->packed.p;
    out->n_head = b->n_head;
    return 0;
}
extern "C" int besst_bam_get_stats(const besst_bam* b, besst_bam_stats* out) {
    if (!b || !out) return -1;
    *out = b->stats;
    return 0;
}
extern "C" void besst_bam_close(besst_bam* b) { delete b; }
