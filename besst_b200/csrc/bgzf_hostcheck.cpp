// libbesst_bgzf_hostcheck.so -- the device BAM ingest's source (bgzf_core.cuh: inflate, CRC-32, record scan, record
// decode; bam_ingest.hpp: window loop and chain verification) compiled for the HOST with a lane-serial warp policy.
//
// TEST TOOLING ONLY: lets the CPU test suite check the code the GPU runs (everything except the four warp primitives and
// the CUDA launch plumbing of besst_bamdev.cu) against zlib and the pure-Python BAM reader on a box without a GPU.  No
// product path loads this library; the product ingest is besst_bam_ingest in libbesst_b200.so (and fails without a GPU).
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <vector>

#include "bam_ingest.hpp"
#include "bgzf_core.cuh"

namespace {

// the warp primitives, one lane after the other: same order of effects as the device versions (literals of a queue first,
// then its matches in queue order with period-`dist` source indexing)
int g_width = 32;

struct HostWarp {
    int w = g_width;   // lanes of the group that works on one block: the device build runs BGZF_GROUP of 32, 16, 8 or 4
    int width() const { return w; }
    bool leader() const { return true; }
    int bcast(int v) const { return v; }
    int place(bgzf::WarpTables* T, int n, uint8_t* out, uint32_t pos, uint32_t usize) const {
        uint32_t P[bgzf::QUEUE], L[bgzf::QUEUE], D[bgzf::QUEUE];
        uint32_t p = pos;
        for (int i = 0; i < n; ++i) {
            D[i] = T->queue[i] >> 16;
            L[i] = D[i] ? (T->queue[i] & 0xffffu) : 1;
            P[i] = p;
            p += L[i];
            if (P[i] + L[i] > usize || D[i] > P[i]) return -1;
        }
        for (int i = 0; i < n; ++i)
            if (!D[i]) out[P[i]] = (uint8_t)T->queue[i];
        for (int i = 0; i < n; ++i)
            if (D[i])
                for (int round = 0; round < (int)L[i]; round += w) {   // one group of lanes at a time, sources read before the stores
                    uint8_t tmp[32];
                    const int m = std::min<int>(w, (int)L[i] - round);
                    for (int l = 0; l < m; ++l) {
                        const uint32_t k = (uint32_t)(round + l);
                        tmp[l] = out[P[i] - D[i] + (D[i] >= L[i] ? k : k % D[i])];
                    }
                    for (int l = 0; l < m; ++l) out[P[i] + round + l] = tmp[l];
                }
        return (int)(p - pos);
    }
    void copy_in(uint8_t* dst, const uint8_t* src, uint32_t n) const { memcpy(dst, src, n); }
};

bgzf::CrcTables g_crc;
bool g_crc_ready = false;
const bgzf::CrcTables* crc_tables() {
    if (!g_crc_ready) {
        bgzf::crc_make_tables(&g_crc);
        g_crc_ready = true;
    }
    return &g_crc;
}

// the 32-lane CRC of the device kernel, lanes in a loop
uint32_t crc_lanes(const uint32_t* words, uint64_t off, uint32_t n) {
    const bgzf::CrcTables* t = crc_tables();
    const uint32_t rounds = n / 128;
    uint32_t folded = 0;
    if (rounds)
        for (int lane = 0; lane < 32; ++lane) folded ^= bgzf::crc_fold_lane(bgzf::crc_lane_rows(t, words, off, rounds, lane), lane);
    uint32_t state = bgzf::crc_with_init(folded, 128ull * rounds);
    state = bgzf::crc_tail(t, state, words, off + 128ull * rounds, n - 128 * rounds);
    return state ^ 0xffffffffu;
}

struct HostBackend {
    int fd = -1;
    int64_t fsize = 0;
    std::string err;
    std::vector<uint32_t> staging[2], ubuf[2];
    std::vector<uint32_t> offs[2];
    std::vector<bamingest::ScanEntry> scan_out[2];
    int32_t inflate_err[2] = {0, 0};
    int64_t inflate_err_block[2] = {-1, -1};
    int64_t crc_bad[2] = {0, 0};
    int64_t decode_bad = 0, unpackable = 0;
    bamingest::Options opt;
    int32_t n_ref = 0;
    bool blind = false;   // seed every hop at the block's first byte: the verification has to repair the wrong ones
    // columns
    std::vector<int32_t> tid, mtid, pos, mpos, tlen, qlen, rlen, alen;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> mapq;
    std::vector<uint32_t> packed;

    int64_t file_size() const { return fsize; }
    std::string error() const { return err; }
    bool load(int buf, int64_t off, int64_t want, const unsigned char** bytes, int64_t* have) {
        const int64_t n = std::min<int64_t>(want, fsize - off);
        staging[buf].assign((size_t)(n + 3) / 4 + 4, 0);
        int64_t got = 0;
        while (got < n) {
            const ssize_t r = pread(fd, reinterpret_cast<unsigned char*>(staging[buf].data()) + got, (size_t)(n - got), (off_t)(off + got));
            if (r <= 0) { err = "read failed"; return false; }
            got += r;
        }
        *bytes = reinterpret_cast<const unsigned char*>(staging[buf].data());
        *have = n;
        return true;
    }
    bool upload(int, const bamingest::Window&) { return true; }
    bool inflate(int buf, const bamingest::Window& W, bool check_crc) {
        const int64_t BASE = opt.carry_max;
        ubuf[buf].resize((size_t)(BASE + W.inflated + 3) / 4 + 4);
        inflate_err[buf] = 0;
        inflate_err_block[buf] = -1;
        crc_bad[buf] = 0;
        bgzf::WarpTables T;
        HostWarp wp;
        uint8_t* out = reinterpret_cast<uint8_t*>(ubuf[buf].data());
        for (size_t k = 0; k < W.blocks.size(); ++k) {
            const bamingest::BlockEntry& b = W.blocks[k];
            const int rc = bgzf::inflate_block(wp, staging[buf].data(), b.cin, b.clen, out + b.out, b.usize, &T);
            if (rc != 0 && inflate_err[buf] == 0) { inflate_err[buf] = rc; inflate_err_block[buf] = (int64_t)k; }
            if (rc == 0 && check_crc && crc_lanes(ubuf[buf].data(), b.out, b.usize) != b.crc) ++crc_bad[buf];
        }
        return true;
    }
    bool read_inflated(int buf, int64_t off, int64_t n, unsigned char* dst) {
        memcpy(dst, reinterpret_cast<const unsigned char*>(ubuf[buf].data()) + off, (size_t)n);
        return true;
    }
    bool scan(int buf, const bamingest::Window& W, int64_t cur, int64_t wend, int32_t nref) {
        n_ref = nref;
        const size_t nb = W.blocks.size();
        offs[buf].assign(nb * bgzf::MAX_RECORDS_PER_BLOCK, 0);
        scan_out[buf].assign(nb, bamingest::ScanEntry{0, 0, 0, 0});
        const uint32_t* u = ubuf[buf].data();
        for (size_t k = 0; k < nb; ++k) {
            const uint64_t b0 = W.blocks[k].out, b1 = b0 + W.blocks[k].usize;
            bamingest::ScanEntry& e = scan_out[buf][k];
            uint64_t seed = ~0ull;
            const bool have_cur = cur >= 0;
            if (have_cur && (((uint64_t)cur >= b0 && (uint64_t)cur < b1) || (k == 0 && (uint64_t)cur < b0))) seed = (uint64_t)cur;
            else if ((!have_cur || (uint64_t)cur < b0) && blind) seed = b0;
            else if (!have_cur || (uint64_t)cur < b0)
                for (uint64_t o = b0; o < b1; ++o)
                    if (bgzf::record_plausible(u, o, (uint64_t)wend, n_ref)) { seed = o; break; }
            if (seed == ~0ull) { e.seed = 0xffffffffu; e.land = 0xffffffffu; e.count = 0; e.flags = 0; continue; }
            e.seed = (uint32_t)seed;
            e.land = (uint32_t)bgzf::hop_block(u, seed, b1, (uint64_t)wend, offs[buf].data() + k * bgzf::MAX_RECORDS_PER_BLOCK, &e.count, &e.flags);
        }
        return true;
    }
    bool inflate_verdict(int buf, std::string* why) {
        if (inflate_err[buf]) { *why = "inflate failed: corrupt deflate stream in BGZF block (code " + std::to_string(inflate_err[buf]) + ")"; return false; }
        if (crc_bad[buf]) { *why = "CRC32 mismatch in " + std::to_string(crc_bad[buf]) + " BGZF block(s)"; return false; }
        return true;
    }
    bool scan_results(int buf, const bamingest::Window&, bamingest::ScanEntry** entries, std::string* why) {
        if (inflate_err[buf]) { *why = "inflate failed: corrupt deflate stream in BGZF block (code " + std::to_string(inflate_err[buf]) + ")"; return false; }
        if (crc_bad[buf]) { *why = "CRC32 mismatch in " + std::to_string(crc_bad[buf]) + " BGZF block(s)"; return false; }
        *entries = scan_out[buf].data();
        return true;
    }
    bool rescan(int buf, const bamingest::Window& W, int64_t k, int64_t start, int64_t wend, bamingest::ScanEntry* e) {
        const uint64_t b1 = (uint64_t)W.blocks[(size_t)k].out + W.blocks[(size_t)k].usize;
        e->seed = (uint32_t)start;
        e->land = (uint32_t)bgzf::hop_block(ubuf[buf].data(), (uint64_t)start, b1, (uint64_t)wend,
                                            offs[buf].data() + (size_t)k * bgzf::MAX_RECORDS_PER_BLOCK, &e->count, &e->flags);
        return true;
    }
    bool decode(int buf, const bamingest::Window& W, const std::vector<bamingest::DecodeEntry>& dec, int64_t, int64_t n_after, int64_t) {
        tid.resize((size_t)n_after); mtid.resize((size_t)n_after); pos.resize((size_t)n_after); mpos.resize((size_t)n_after);
        tlen.resize((size_t)n_after); qlen.resize((size_t)n_after); flag.resize((size_t)n_after); mapq.resize((size_t)n_after);
        packed.resize((size_t)n_after);
        const size_t nh = (size_t)std::min<int64_t>(n_after, opt.head_records);
        rlen.resize(nh); alen.resize(nh);
        const uint32_t* u = ubuf[buf].data();
        for (size_t k = 0; k < W.blocks.size(); ++k)
            for (uint32_t i = 0; i < dec[k].count; ++i) {
                const bgzf::RecordFields f = bgzf::decode_record(u, offs[buf][k * bgzf::MAX_RECORDS_PER_BLOCK + i]);
                const size_t g = (size_t)(dec[k].base + i);
                if (!f.ok) ++decode_bad;
                tid[g] = f.tid; mtid[g] = f.mtid; pos[g] = f.pos; mpos[g] = f.mpos; tlen[g] = f.tlen; qlen[g] = f.qlen;
                flag[g] = (uint16_t)f.flag; mapq[g] = (uint8_t)f.mapq;
                if (f.flag >= 4096u || f.qlen < 0 || f.qlen >= 4096) ++unpackable;
                packed[g] = (f.flag & 0xfffu) | (f.mapq << 12) | ((uint32_t)f.qlen << 20);
                if (g < nh) { rlen[g] = f.rlen; alen[g] = f.alen; }
            }
        return true;
    }
    bool carry(int from, int64_t src, int64_t n, int to, int64_t dst) {
        const int64_t BASE = opt.carry_max;
        if (ubuf[to].size() * 4 < (size_t)BASE + 16) ubuf[to].resize((size_t)BASE / 4 + 4);
        memmove(reinterpret_cast<unsigned char*>(ubuf[to].data()) + dst, reinterpret_cast<const unsigned char*>(ubuf[from].data()) + src, (size_t)n);
        return true;
    }
    bool finish(std::string*, bool* bad_records) {
        *bad_records = decode_bad != 0;
        decode_bad = 0;
        return true;
    }
};

struct Handle {
    HostBackend B;
    bamingest::Result res;
    std::string err;
};

}  // namespace

extern "C" {

// lanes per block of the emulated placement rounds (the device build's BGZF_GROUP)
void bgzf_hc_set_width(int w) { g_width = (w == 4 || w == 8 || w == 16) ? w : 32; }

// raw-deflate stream -> usize bytes, through the device decoder's code path; `misalign` (0..3) shifts the stream inside the
// word buffer.  -> 0 or a bgzf::E_* code
int bgzf_hc_inflate(const uint8_t* cdata, uint32_t clen, int misalign, uint8_t* out, uint32_t usize) {
    std::vector<uint32_t> words((size_t)(clen + misalign + 3) / 4 + 4, 0);
    memcpy(reinterpret_cast<uint8_t*>(words.data()) + misalign, cdata, clen);
    bgzf::WarpTables T;
    HostWarp wp;
    return bgzf::inflate_block(wp, words.data(), (uint64_t)misalign, clen, out, usize, &T);
}

uint32_t bgzf_hc_crc32(const uint8_t* data, uint32_t n, int misalign) {
    std::vector<uint32_t> words((size_t)(n + misalign + 3) / 4 + 4, 0);
    memcpy(reinterpret_cast<uint8_t*>(words.data()) + misalign, data, n);
    return crc_lanes(words.data(), (uint64_t)misalign, n);
}

void* bgzf_hc_ingest_part(const char* path, int64_t window_bytes, int64_t max_inflated, int64_t carry_max, int check_crc, int blind,
                          int64_t head_records, int part, int n_parts, int64_t start_voffset, int64_t tail_bytes, char* err, int err_len);

void* bgzf_hc_ingest(const char* path, int64_t window_bytes, int64_t max_inflated, int64_t carry_max, int check_crc, int blind, int64_t head_records,
                     char* err, int err_len) {
    return bgzf_hc_ingest_part(path, window_bytes, max_inflated, carry_max, check_crc, blind, head_records, 0, 1, -1, 0, err, err_len);
}

// one part of the file (the multi-GPU ingest's per-rank call): see bamingest::Options
void* bgzf_hc_ingest_part(const char* path, int64_t window_bytes, int64_t max_inflated, int64_t carry_max, int check_crc, int blind,
                          int64_t head_records, int part, int n_parts, int64_t start_voffset, int64_t tail_bytes, char* err, int err_len) {
    Handle* h = new Handle();
    auto fail = [&](const std::string& m) -> void* {
        if (err && err_len > 0) { strncpy(err, m.c_str(), (size_t)err_len - 1); err[err_len - 1] = 0; }
        if (h->B.fd >= 0) close(h->B.fd);
        delete h;
        return nullptr;
    };
    h->B.blind = blind != 0;
    h->B.fd = open(path, O_RDONLY);
    if (h->B.fd < 0) return fail(std::string("cannot open ") + path);
    struct stat st;
    if (fstat(h->B.fd, &st) != 0) return fail("cannot stat");
    h->B.fsize = (int64_t)st.st_size;
    bamingest::Options opt;
    if (window_bytes > 0) opt.window_bytes = window_bytes;
    if (max_inflated > 0) opt.max_inflated = max_inflated;
    if (carry_max > 0) opt.carry_max = carry_max;
    opt.check_crc = check_crc != 0;
    opt.head_records = head_records;
    opt.part = part;
    opt.n_parts = n_parts;
    opt.start_voffset = start_voffset;
    if (tail_bytes > 0) opt.tail_bytes = tail_bytes;
    std::string why;
    for (;;) {
        h->B.opt = opt;
        const int rc = bamingest::run(h->B, opt, &h->res, &why);
        if (rc == bamingest::RC_WINDOW_TOO_SMALL) {   // the header did not fit into the first window
            opt.window_bytes *= 4;
            opt.max_inflated = std::max(opt.max_inflated, opt.window_bytes * 8);
            continue;
        }
        if (rc != bamingest::RC_OK) return fail(why);
        break;
    }
    close(h->B.fd);
    h->B.fd = -1;
    return h;
}
int64_t bgzf_hc_n(void* h) { return static_cast<Handle*>(h)->res.n_records; }
int64_t bgzf_hc_n_head(void* h) { return static_cast<Handle*>(h)->res.n_head; }
int64_t bgzf_hc_n_refs(void* h) { return (int64_t)static_cast<Handle*>(h)->res.ref_names.size(); }
const char* bgzf_hc_ref_name(void* h, int64_t i) { return static_cast<Handle*>(h)->res.ref_names[(size_t)i].c_str(); }
int64_t bgzf_hc_ref_length(void* h, int64_t i) { return static_cast<Handle*>(h)->res.ref_lengths[(size_t)i]; }
int64_t bgzf_hc_stat(void* h, int which) {
    const bamingest::Result& r = static_cast<Handle*>(h)->res;
    const bamingest::Stats& s = r.stats;
    const int64_t v[] = {s.compressed_bytes, s.uncompressed_bytes, s.blocks, s.records, s.windows, s.rescans, static_cast<Handle*>(h)->B.unpackable,
                         r.first_voffset, r.landing_voffset};
    return which >= 0 && which < 9 ? v[which] : -1;
}
const void* bgzf_hc_column(void* hp, int which) {
    HostBackend& B = static_cast<Handle*>(hp)->B;
    switch (which) {
        case 0: return B.tid.data();
        case 1: return B.mtid.data();
        case 2: return B.pos.data();
        case 3: return B.mpos.data();
        case 4: return B.tlen.data();
        case 5: return B.qlen.data();
        case 6: return B.flag.data();
        case 7: return B.mapq.data();
        case 8: return B.packed.data();
        case 9: return B.rlen.data();
        case 10: return B.alen.data();
    }
    return nullptr;
}
void bgzf_hc_close(void* h) { delete static_cast<Handle*>(h); }

}  // extern "C"
