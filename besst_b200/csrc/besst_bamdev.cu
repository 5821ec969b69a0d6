// besst_bamdev.cu -- BAM file -> record columns resident in HBM, inflated and decoded ON THE GPU.
//
// Replaces the per-record pysam iteration in front of the hot path (runBESST:162, libmetrics.py:63,257,293,
// CreateGraph.py:111).  The file crosses PCIe COMPRESSED; every BGZF block is one warp's raw-deflate job
// (bgzf_core.cuh), its CRC-32 is checked by the same warp, record boundaries are found per block and verified on the
// host in O(blocks) (bam_ingest.hpp), and the fixed-core fields land in the column layout of besst_records -- already
// on the device, so besst_libmetrics / besst_graph_build run on them with on_device = 1 and no record ever visits host
// memory.  Windows of the file are double-buffered: the host read + upload of window w+1 overlaps the kernels of window w.
//
// Kernels (sm_100a; all HBM/latency-bound byte work, no tensor-core shape):
//   k_bgzf_inflate   8 warps per CTA, one BGZF block per warp; 3.6 KB of decode tables per warp + 5 KB CRC tables per CTA in
//                    shared memory; the leader lane decodes 32 symbols, the warp places them (shuffle scan, coalesced
//                    literal store, warp-wide match copies), then the 32-lane CRC over the block it just wrote (L2-hot)
//   k_bam_scan       one warp per block: 32 candidate record starts per step (ballot), leader hops the block_size chain
//   k_bam_rescan     one thread: re-hop of a block whose seed the host verification rejected
//   k_bam_decode     one warp per block, one record per lane: unaligned field loads by funnel shift, coalesced column stores
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <thread>

#include "bam_ingest.hpp"
#include "besst_internal.cuh"
#include "bgzf_core.cuh"

namespace {

// lanes that work on one BGZF block (tuning knob): with BGZF_GROUP < 32 a warp decodes 32 / BGZF_GROUP blocks at once, the
// leader lanes of the groups running the serial symbol decode side by side and the placement rounds / CRC using group-wide
// shuffles.  The kernel is issue-bound (ncu, one block per warp: 71 % SM throughput, 2.85 IPC, 18.4 G warp instructions for
// 618 MB of output), but sharing issue slots between leaders does NOT pay: measured on a 4.04 M-record file (832 MB
// inflated), inflate time 30.9 ms with 32 lanes per block, 38.8 / 53.0 ms with 16 / 8 (fewer warps per SM for the same
// shared-memory footprint: latency-bound), 31.6 / 38.3 / 58.4 ms with 16 / 8 / 4 and the small tables (2^9 / 2^7 entries,
// 16 blocks per CTA) -- profiles/r02/ingest_variants.md.  One warp per block stays.
#ifndef BGZF_GROUP
#define BGZF_GROUP 32
#endif
constexpr int GROUP = BGZF_GROUP;
#ifndef BGZF_BLOCKS_PER_CTA
#define BGZF_BLOCKS_PER_CTA 8
#endif
constexpr int INFLATE_BLOCKS_PER_CTA = BGZF_BLOCKS_PER_CTA;     // BGZF blocks per CTA (3.6 KB of tables each, 5 KB of CRC tables per CTA)
constexpr int INFLATE_THREADS = INFLATE_BLOCKS_PER_CTA * GROUP;
constexpr int E_CRC = -10;
static_assert(GROUP == 32 || GROUP == 16 || GROUP == 8 || GROUP == 4, "BGZF_GROUP");
static_assert(INFLATE_THREADS % 32 == 0, "whole warps");

// ---- the warp primitives of bgzf::inflate_block on the device -----------------------------------------------------------
struct DevWarp {
    int gl;          // lane inside the group
    unsigned gmask;  // the group's lanes inside the warp
    int gshift;      // first lane of the group
    __device__ __forceinline__ int width() const { return GROUP; }
    __device__ __forceinline__ bool leader() const { return gl == 0; }
    __device__ __forceinline__ int bcast(int v) const {
        __syncwarp(gmask);
        return __shfl_sync(gmask, v, 0, GROUP);
    }
    // the queue's n symbols -> out[pos ..): literals first (one store per lane), then the matches in queue order, each a
    // group-wide copy.  A match may read what an earlier symbol of the same queue wrote (literals are all in place, earlier
    // matches are complete); with dist < len the source repeats with period dist, all of it in front of the match.
    __device__ __forceinline__ int place(bgzf::WarpTables* T, int n, uint8_t* out, uint32_t pos, uint32_t usize) const {
        uint32_t D = 0, L = 0, lit = 0;
        if (gl < n) {
            const uint32_t q = T->queue[gl];
            D = q >> 16;
            lit = q & 0xffffu;
            L = D ? lit : 1u;
        }
        uint32_t x = L;
#pragma unroll
        for (int o = 1; o < GROUP; o <<= 1) {
            const uint32_t y = __shfl_up_sync(gmask, x, o, GROUP);
            if (gl >= o) x += y;
        }
        const uint32_t P = pos + x - L;
        const uint32_t total = __shfl_sync(gmask, x, GROUP - 1, GROUP);
        const bool bad = gl < n && (P + L > usize || D > P);
        if (__any_sync(gmask, bad)) return -1;
        if (gl < n && D == 0) out[P] = (uint8_t)lit;
        __syncwarp(gmask);
        unsigned m = (__ballot_sync(gmask, gl < n && D != 0) & gmask) >> gshift;
        while (m) {
            const int s = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t Ps = __shfl_sync(gmask, P, s, GROUP), Ls = __shfl_sync(gmask, L, s, GROUP), Ds = __shfl_sync(gmask, D, s, GROUP);
            const uint8_t* src = out + Ps - Ds;
            if (Ds >= Ls) {
                for (uint32_t i = gl; i < Ls; i += GROUP) out[Ps + i] = src[i];
            } else {
                for (uint32_t i = gl; i < Ls; i += GROUP) out[Ps + i] = src[i % Ds];
            }
            __syncwarp(gmask);
        }
        return (int)total;
    }
    __device__ __forceinline__ void copy_in(uint8_t* dst, const uint8_t* src, uint32_t n) const {
        for (uint32_t i = gl; i < n; i += GROUP) dst[i] = src[i];
        __syncwarp(gmask);
    }
};

// err[0] blocks whose inflate failed, err[1] first such block (min), err[2] its code, err[3] CRC mismatches,
// err[4] records whose name/CIGAR overrun the record, err[5] records that do not fit the packed column
__global__ void __launch_bounds__(INFLATE_THREADS)
k_bgzf_inflate(const uint32_t* __restrict__ cwords, const bamingest::BlockEntry* __restrict__ blocks, int n_blocks, uint8_t* ubuf,
               const bgzf::CrcTables* __restrict__ crc_tab, int check_crc, int* err) {
    __shared__ bgzf::WarpTables tables[INFLATE_BLOCKS_PER_CTA];
    __shared__ bgzf::CrcTables crc;
    if (check_crc) {
        const uint32_t* s = reinterpret_cast<const uint32_t*>(crc_tab);
        uint32_t* d = reinterpret_cast<uint32_t*>(&crc);
        for (int i = threadIdx.x; i < (int)(sizeof(bgzf::CrcTables) / 4); i += blockDim.x) d[i] = s[i];
        __syncthreads();
    }
    const int slot = threadIdx.x / GROUP;   // which of the CTA's blocks
    const int k = blockIdx.x * INFLATE_BLOCKS_PER_CTA + slot;
    if (k >= n_blocks) return;              // whole groups leave; the others only ever synchronise inside their group
    const int lane = threadIdx.x & 31;
    DevWarp wp;
    wp.gl = lane % GROUP;
    wp.gshift = lane - wp.gl;
    wp.gmask = (GROUP == 32 ? 0xffffffffu : ((1u << (GROUP & 31)) - 1u)) << wp.gshift;
    const bamingest::BlockEntry b = blocks[k];
    int rc = bgzf::inflate_block(wp, cwords, (uint64_t)b.cin, b.clen, ubuf + b.out, b.usize, &tables[slot]);
    __syncwarp(wp.gmask);
    if (rc == 0 && check_crc) {
        // 32 virtual lanes (one per word of a 128-byte row), 32 / GROUP of them per lane
        const uint32_t* u = reinterpret_cast<const uint32_t*>(ubuf);
        const uint32_t rounds = b.usize / 128;
        uint32_t folded = 0;
        if (rounds) {
#pragma unroll
            for (int v = 0; v < 32 / GROUP; ++v) {
                const int vl = wp.gl + v * GROUP;
                folded ^= bgzf::crc_fold_lane(bgzf::crc_lane_rows(&crc, u, b.out, rounds, vl), vl);
            }
#pragma unroll
            for (int o = GROUP / 2; o; o >>= 1) folded ^= __shfl_xor_sync(wp.gmask, folded, o, GROUP);
        }
        if (wp.gl == 0) {
            uint32_t st = bgzf::crc_with_init(folded, 128ull * rounds);
            st = bgzf::crc_tail(&crc, st, u, b.out + 128ull * rounds, b.usize - 128 * rounds) ^ 0xffffffffu;
            if (st != b.crc) {
                atomicAdd(&err[3], 1);
                rc = E_CRC;
            }
        }
    }
    if (wp.gl == 0 && rc != 0 && rc != E_CRC) {
        atomicAdd(&err[0], 1);
        if (atomicMin(&err[1], k) > k) err[2] = rc;
    }
}

__global__ void __launch_bounds__(256)
k_bam_scan(const uint32_t* __restrict__ u, const bamingest::BlockEntry* __restrict__ blocks, int n_blocks, int have_cur, uint64_t cur,
           uint64_t wend, int32_t n_ref, int blind, uint32_t* __restrict__ offs, bamingest::ScanEntry* __restrict__ out) {
    const int k = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (k >= n_blocks) return;
    const uint64_t b0 = blocks[k].out, b1 = b0 + blocks[k].usize;
    uint64_t seed = ~0ull;
    if (have_cur && ((cur >= b0 && cur < b1) || (k == 0 && cur < b0))) {
        seed = cur;
    } else if (!have_cur || cur < b0) {   // no known start (a part behind the header): every block guesses
        if (blind) {
            seed = b0;
        } else {
            for (uint64_t base = b0; base < b1; base += 32) {
                const uint64_t o = base + lane;
                const bool ok = o < b1 && bgzf::record_plausible(u, o, wend, n_ref);
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (m) {
                    seed = base + (uint64_t)(__ffs(m) - 1);
                    break;
                }
            }
        }
    }
    if (lane != 0) return;
    bamingest::ScanEntry e;
    if (seed == ~0ull) {
        e.seed = 0xffffffffu; e.land = 0xffffffffu; e.count = 0; e.flags = 0;
    } else {
        e.seed = (uint32_t)seed;
        e.land = (uint32_t)bgzf::hop_block(u, seed, b1, wend, offs + (size_t)k * bgzf::MAX_RECORDS_PER_BLOCK, &e.count, &e.flags);
    }
    out[k] = e;
}

__global__ void k_bam_rescan(const uint32_t* __restrict__ u, const bamingest::BlockEntry* __restrict__ blocks, int k, uint64_t start,
                             uint64_t wend, uint32_t* __restrict__ offs, bamingest::ScanEntry* __restrict__ out) {
    const uint64_t b1 = (uint64_t)blocks[k].out + blocks[k].usize;
    bamingest::ScanEntry e;
    e.seed = (uint32_t)start;
    e.land = (uint32_t)bgzf::hop_block(u, start, b1, wend, offs + (size_t)k * bgzf::MAX_RECORDS_PER_BLOCK, &e.count, &e.flags);
    out[k] = e;
}

struct Columns {
    int32_t *tid, *mtid, *pos, *mpos, *tlen, *qlen;
    uint16_t* flag;
    uint8_t* mapq;
    uint32_t* packed;
    int32_t *rlen, *alen;   // first n_head records
};

__global__ void __launch_bounds__(256)
k_bam_decode(const uint32_t* __restrict__ u, const uint32_t* __restrict__ offs, const bamingest::DecodeEntry* __restrict__ dec, int n_blocks,
             Columns c, int64_t n_head, int* err) {
    const int k = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (k >= n_blocks) return;
    const uint32_t n = dec[k].count;
    const int64_t base = dec[k].base;
    const uint32_t* my = offs + (size_t)k * bgzf::MAX_RECORDS_PER_BLOCK;
    for (uint32_t i = lane; i < n; i += 32) {
        const bgzf::RecordFields f = bgzf::decode_record(u, my[i]);
        const int64_t g = base + i;
        c.tid[g] = f.tid; c.mtid[g] = f.mtid; c.pos[g] = f.pos; c.mpos[g] = f.mpos; c.tlen[g] = f.tlen; c.qlen[g] = f.qlen;
        c.flag[g] = (uint16_t)f.flag;
        c.mapq[g] = (uint8_t)f.mapq;
        c.packed[g] = (f.flag & 0xfffu) | (f.mapq << 12) | ((uint32_t)f.qlen << 20);
        if (!f.ok) atomicAdd(&err[4], 1);
        if (f.flag >= 4096u || f.qlen < 0 || f.qlen >= 4096) atomicAdd(&err[5], 1);
        if (g < n_head) { c.rlen[g] = f.rlen; c.alen[g] = f.alen; }
    }
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

// ---- state kept in the ctx between besst_bam_ingest and the calls that consume its columns -----------------------------------
struct BesstBamIngest {
    DBuf col_i32[6], col_flag, col_mapq, col_packed, head_rlen, head_alen;
    int64_t cap = 0, n = 0, n_head = 0;
    bool unpackable = false;
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lengths;
    // window machinery (kept for the next file)
    HBuf staging[2], tbl_h[2], scan_h[2], dec_h[2], err_h;
    DBuf cbuf[2], tbl_d[2], ubuf[2], offs[2], scan_d[2], dec_d[2], err_d, crc_d;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_inflated[2] = {nullptr, nullptr}, ev_scan[2] = {nullptr, nullptr};
    bool h2d_pending[2] = {false, false}, inflated_pending[2] = {false, false};
    bool crc_ready = false;
    besst_bam_ingest_stats stats;
};

void besst_bamdev_release(besst_ctx* ctx) {
    BesstBamIngest* s = ctx->ingest;
    if (!s) return;
    for (DBuf& b : s->col_i32) b.release();
    DBuf* d[] = {&s->col_flag, &s->col_mapq, &s->col_packed, &s->head_rlen, &s->head_alen, &s->err_d, &s->crc_d};
    for (DBuf* b : d) b->release();
    for (int i = 0; i < 2; ++i) {
        s->staging[i].release(); s->tbl_h[i].release(); s->scan_h[i].release(); s->dec_h[i].release();
        s->cbuf[i].release(); s->tbl_d[i].release(); s->ubuf[i].release(); s->offs[i].release(); s->scan_d[i].release(); s->dec_d[i].release();
        if (s->ev_h2d[i]) cudaEventDestroy(s->ev_h2d[i]);
        if (s->ev_inflated[i]) cudaEventDestroy(s->ev_inflated[i]);
        if (s->ev_scan[i]) cudaEventDestroy(s->ev_scan[i]);
    }
    s->err_h.release();
    delete s;
    ctx->ingest = nullptr;
}

namespace {

// the window loop's Backend on CUDA streams: kernels on ctx->stream, uploads on ctx->copy_stream
struct DevBackend {
    besst_ctx* ctx;
    BesstBamIngest* S;
    bamingest::Options opt;
    int fd = -1;
    int64_t fsize = 0;
    int flags = 0;
    std::string err;
    int read_threads = 4;
    int64_t chunk_bytes = 32ll << 20;   // file bytes per read + upload step
    double t_read = 0;
    struct Ev { cudaEvent_t a, b; int what; };
    std::vector<Ev> timers;
    int64_t inflated_win[2] = {0, 0};   // inflated bytes of the window uploaded into each buffer pair

    bool ok(cudaError_t e, const char* what) {
        if (e == cudaSuccess) return true;
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return false;
    }
    void tick(int what, cudaEvent_t* a) {
        Ev e; e.what = what;
        cudaEventCreate(&e.a); cudaEventCreate(&e.b);
        cudaEventRecord(e.a, ctx->stream);
        timers.push_back(e);
        *a = e.b;
    }
    int64_t file_size() const { return fsize; }
    std::string error() const { return err; }

    // file -> pinned staging -> device, chunk by chunk: the host threads read chunk c + 1 while chunk c crosses PCIe
    bool load(int buf, int64_t off, int64_t want, const unsigned char** bytes, int64_t* have) {
        const double t0 = now_s();
        if (S->h2d_pending[buf]) {   // the previous upload out of this staging buffer
            if (!ok(cudaEventSynchronize(S->ev_h2d[buf]), "cudaEventSynchronize(upload)")) return false;
            S->h2d_pending[buf] = false;
        }
        const int64_t n = std::min<int64_t>(want, fsize - off);
        if (!ok(S->staging[buf].ensure((size_t)n + 64), "cudaHostAlloc(staging)")) return false;
        if (S->inflated_pending[buf]) {   // the kernel that still reads this compressed buffer / block table
            if (!ok(cudaStreamWaitEvent(ctx->copy_stream, S->ev_inflated[buf], 0), "cudaStreamWaitEvent")) return false;
        }
        // growing a buffer frees the old one: cudaFree synchronises, so a kernel still reading it has finished
        if (!ok(S->cbuf[buf].ensure((size_t)n + 64), "cudaMalloc(compressed window)")) return false;
        unsigned char* dst = static_cast<unsigned char*>(S->staging[buf].p);
        memset(dst + n, 0, 64);
        for (int64_t c0 = 0; c0 < n || c0 == 0; c0 += chunk_bytes) {
            const int64_t c1 = std::min<int64_t>(n, c0 + chunk_bytes), m = c1 - c0;
            const int nt = m >= (8 << 20) ? read_threads : 1;
            std::vector<std::thread> th;
            std::vector<int> bad((size_t)nt, 0);
            auto work = [&](int t) {
                int64_t a = c0 + m * t / nt;
                const int64_t b = c0 + m * (t + 1) / nt;
                while (a < b) {
                    const ssize_t r = pread(fd, dst + a, (size_t)(b - a), (off_t)(off + a));
                    if (r <= 0) { bad[(size_t)t] = 1; return; }
                    a += r;
                }
            };
            for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
            work(0);
            for (auto& x : th) x.join();
            for (int v : bad) if (v) { err = "read failed"; return false; }
            const int64_t bytes_now = m + (c1 == n ? 64 : 0);
            if (!ok(cudaMemcpyAsync(S->cbuf[buf].as<unsigned char>() + c0, dst + c0, (size_t)bytes_now, cudaMemcpyHostToDevice, ctx->copy_stream), "H2D window"))
                return false;
            if (c1 >= n) break;
        }
        S->h2d_pending[buf] = true;   // staging is in flight from here on; upload() records the event
        if (!ok(cudaEventRecord(S->ev_h2d[buf], ctx->copy_stream), "cudaEventRecord")) return false;
        *bytes = dst;
        *have = n;
        t_read += now_s() - t0;
        return true;
    }

    bool upload(int buf, const bamingest::Window& W) {
        const size_t nb = W.blocks.size();
        if (!ok(S->tbl_h[buf].ensure(nb * sizeof(bamingest::BlockEntry)), "cudaHostAlloc(block table)")) return false;
        memcpy(S->tbl_h[buf].p, W.blocks.data(), nb * sizeof(bamingest::BlockEntry));
        if (!ok(S->tbl_d[buf].ensure(nb * sizeof(bamingest::BlockEntry)), "cudaMalloc(block table)")) return false;
        if (!ok(cudaMemcpyAsync(S->tbl_d[buf].p, S->tbl_h[buf].p, nb * sizeof(bamingest::BlockEntry), cudaMemcpyHostToDevice, ctx->copy_stream), "H2D block table")) return false;
        if (!ok(cudaEventRecord(S->ev_h2d[buf], ctx->copy_stream), "cudaEventRecord")) return false;
        S->h2d_pending[buf] = true;
        inflated_win[buf] = W.inflated;
        return true;
    }

    bool inflate(int buf, const bamingest::Window& W, bool check_crc) {
        const int nb = (int)W.blocks.size();
        if (!ok(S->ubuf[buf].ensure((size_t)(opt.carry_max + W.inflated) + 64), "cudaMalloc(inflated window)")) return false;
        if (!ok(cudaStreamWaitEvent(ctx->stream, S->ev_h2d[buf], 0), "cudaStreamWaitEvent")) return false;
        cudaEvent_t done;
        tick(0, &done);
        {
            KTimer kt(ctx, BESST_K_BAM_INFLATE);
            k_bgzf_inflate<<<(nb + INFLATE_BLOCKS_PER_CTA - 1) / INFLATE_BLOCKS_PER_CTA, INFLATE_THREADS, 0, ctx->stream>>>(
                S->cbuf[buf].as<uint32_t>(), S->tbl_d[buf].as<bamingest::BlockEntry>(), nb, S->ubuf[buf].as<uint8_t>(),
                S->crc_d.as<bgzf::CrcTables>(), check_crc ? 1 : 0, S->err_d.as<int>());
        }
        cudaEventRecord(done, ctx->stream);
        if (!ok(cudaEventRecord(S->ev_inflated[buf], ctx->stream), "cudaEventRecord")) return false;
        S->inflated_pending[buf] = true;
        return ok(cudaGetLastError(), "k_bgzf_inflate");
    }

    bool read_inflated(int buf, int64_t off, int64_t n, unsigned char* dst) {
        if (!ok(cudaMemcpyAsync(dst, S->ubuf[buf].as<uint8_t>() + off, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream), "D2H header")) return false;
        return ok(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    }

    bool scan(int buf, const bamingest::Window& W, int64_t cur, int64_t wend, int32_t n_ref) {
        const int nb = (int)W.blocks.size();
        if (!ok(S->offs[buf].ensure((size_t)nb * bgzf::MAX_RECORDS_PER_BLOCK * 4), "cudaMalloc(record offsets)")) return false;
        if (!ok(S->scan_d[buf].ensure((size_t)nb * sizeof(bamingest::ScanEntry)), "cudaMalloc(scan)")) return false;
        if (!ok(S->scan_h[buf].ensure((size_t)nb * sizeof(bamingest::ScanEntry)), "cudaHostAlloc(scan)")) return false;
        cudaEvent_t done;
        tick(1, &done);
        {
            KTimer kt(ctx, BESST_K_BAM_SCAN);
            k_bam_scan<<<(nb + 7) / 8, 256, 0, ctx->stream>>>(S->ubuf[buf].as<uint32_t>(), S->tbl_d[buf].as<bamingest::BlockEntry>(), nb,
                                                              cur >= 0 ? 1 : 0, (uint64_t)(cur >= 0 ? cur : 0), (uint64_t)wend, n_ref,
                                                              (flags & BESST_BAM_BLIND_SEEDS) ? 1 : 0,
                                                              S->offs[buf].as<uint32_t>(), S->scan_d[buf].as<bamingest::ScanEntry>());
        }
        cudaEventRecord(done, ctx->stream);
        if (!ok(cudaGetLastError(), "k_bam_scan")) return false;
        if (!ok(cudaMemcpyAsync(S->scan_h[buf].p, S->scan_d[buf].p, (size_t)nb * sizeof(bamingest::ScanEntry), cudaMemcpyDeviceToHost, ctx->stream), "D2H scan")) return false;
        if (!ok(cudaMemcpyAsync(S->err_h.p, S->err_d.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "D2H status")) return false;
        return ok(cudaEventRecord(S->ev_scan[buf], ctx->stream), "cudaEventRecord");
    }

    bool verdict(std::string* why) {
        const int* e = static_cast<const int*>(S->err_h.p);
        if (e[0]) {
            *why = "inflate failed on the device: corrupt deflate stream in " + std::to_string(e[0]) + " BGZF block(s) (first: block " +
                   std::to_string(e[1]) + " of its window, code " + std::to_string(e[2]) + ")";
            return false;
        }
        if (e[3]) { *why = "CRC32 mismatch in " + std::to_string(e[3]) + " BGZF block(s)"; return false; }
        return true;
    }
    bool inflate_verdict(int, std::string* why) {   // a part's header window: no scan follows it
        if (!ok(cudaMemcpyAsync(S->err_h.p, S->err_d.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "D2H status") ||
            !ok(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize")) { *why = err; return false; }
        return verdict(why);
    }

    bool scan_results(int buf, const bamingest::Window&, bamingest::ScanEntry** entries, std::string* why) {
        if (!ok(cudaEventSynchronize(S->ev_scan[buf]), "cudaEventSynchronize(scan)")) { *why = err; return false; }
        if (!verdict(why)) return false;
        *entries = static_cast<bamingest::ScanEntry*>(S->scan_h[buf].p);
        return true;
    }
    bool rescan(int buf, const bamingest::Window&, int64_t k, int64_t start, int64_t wend, bamingest::ScanEntry* e) {
        {
            KTimer kt(ctx, BESST_K_BAM_SCAN);
            k_bam_rescan<<<1, 1, 0, ctx->stream>>>(S->ubuf[buf].as<uint32_t>(), S->tbl_d[buf].as<bamingest::BlockEntry>(), (int)k, (uint64_t)start,
                                                   (uint64_t)wend, S->offs[buf].as<uint32_t>(), S->scan_d[buf].as<bamingest::ScanEntry>());
        }
        bamingest::ScanEntry* h = static_cast<bamingest::ScanEntry*>(S->scan_h[buf].p) + k;
        if (!ok(cudaMemcpyAsync(h, S->scan_d[buf].as<bamingest::ScanEntry>() + k, sizeof(bamingest::ScanEntry), cudaMemcpyDeviceToHost, ctx->stream), "D2H rescan")) return false;
        if (!ok(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize")) return false;
        *e = *h;
        return true;
    }

    // grow the columns, keeping the first `keep` records (the copy is ordered on the stream)
    bool grow(int64_t need, int64_t keep, int64_t hint) {
        if (need <= S->cap) return true;
        int64_t want = std::max<int64_t>(need, hint);
        if (S->cap) want = std::max<int64_t>(want, S->cap + S->cap / 2);
        want = (want + 1023) / 1024 * 1024;
        auto regrow = [&](DBuf& b, size_t elem) -> bool {
            void* p = nullptr;
            if (!ok(cudaMalloc(&p, (size_t)want * elem), "cudaMalloc(record columns)")) return false;
            if (keep > 0 && b.p && !ok(cudaMemcpyAsync(p, b.p, (size_t)keep * elem, cudaMemcpyDeviceToDevice, ctx->stream), "D2D grow")) return false;
            if (b.p) {
                if (!ok(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize")) return false;
                cudaFree(b.p);
            }
            b.p = p;
            b.cap = (size_t)want * elem;
            return true;
        };
        for (DBuf& b : S->col_i32) if (!regrow(b, 4)) return false;
        if (!regrow(S->col_flag, 2) || !regrow(S->col_mapq, 1) || !regrow(S->col_packed, 4)) return false;
        S->cap = want;
        return true;
    }

    bool decode(int buf, const bamingest::Window& W, const std::vector<bamingest::DecodeEntry>& dec, int64_t n_before, int64_t n_after, int64_t est_total) {
        const int nb = (int)W.blocks.size();
        if (!grow(n_after, n_before, est_total)) return false;
        if (!ok(S->dec_h[buf].ensure((size_t)nb * sizeof(bamingest::DecodeEntry)), "cudaHostAlloc(decode table)")) return false;
        if (!ok(S->dec_d[buf].ensure((size_t)nb * sizeof(bamingest::DecodeEntry)), "cudaMalloc(decode table)")) return false;
        memcpy(S->dec_h[buf].p, dec.data(), (size_t)nb * sizeof(bamingest::DecodeEntry));
        if (!ok(cudaMemcpyAsync(S->dec_d[buf].p, S->dec_h[buf].p, (size_t)nb * sizeof(bamingest::DecodeEntry), cudaMemcpyHostToDevice, ctx->stream), "H2D decode table")) return false;
        Columns c;
        c.tid = S->col_i32[0].as<int32_t>(); c.mtid = S->col_i32[1].as<int32_t>(); c.pos = S->col_i32[2].as<int32_t>();
        c.mpos = S->col_i32[3].as<int32_t>(); c.tlen = S->col_i32[4].as<int32_t>(); c.qlen = S->col_i32[5].as<int32_t>();
        c.flag = S->col_flag.as<uint16_t>(); c.mapq = S->col_mapq.as<uint8_t>(); c.packed = S->col_packed.as<uint32_t>();
        c.rlen = S->head_rlen.as<int32_t>(); c.alen = S->head_alen.as<int32_t>();
        cudaEvent_t done;
        tick(2, &done);
        {
            KTimer kt(ctx, BESST_K_BAM_DECODE);
            k_bam_decode<<<(nb + 7) / 8, 256, 0, ctx->stream>>>(S->ubuf[buf].as<uint32_t>(), S->offs[buf].as<uint32_t>(), S->dec_d[buf].as<bamingest::DecodeEntry>(),
                                                                nb, c, opt.head_records, S->err_d.as<int>());
        }
        cudaEventRecord(done, ctx->stream);
        return ok(cudaGetLastError(), "k_bam_decode");
    }

    bool carry(int from, int64_t src, int64_t n, int to, int64_t dst) {
        // full size of the destination window now: growing the buffer later would drop the carried bytes
        if (!ok(S->ubuf[to].ensure((size_t)(opt.carry_max + inflated_win[to]) + 64), "cudaMalloc(inflated window)")) return false;
        return ok(cudaMemcpyAsync(S->ubuf[to].as<uint8_t>() + dst, S->ubuf[from].as<uint8_t>() + src, (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream), "D2D carry");
    }

    bool finish(std::string* why, bool* bad_records) {
        if (!ok(cudaMemcpyAsync(S->err_h.p, S->err_d.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "D2H status") ||
            !ok(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize") || !ok(cudaStreamSynchronize(ctx->copy_stream), "cudaStreamSynchronize")) {
            *why = err;
            return false;
        }
        const int* e = static_cast<const int*>(S->err_h.p);
        *bad_records = e[4] != 0;
        S->unpackable = e[5] != 0;
        return true;
    }
};

}  // namespace

extern "C" int besst_bam_ingest(besst_ctx* ctx, const char* path, int64_t head_records, int32_t flags, besst_records* out,
                                besst_bam_ingest_stats* stats) {
    return besst_bam_ingest_part(ctx, path, head_records, flags, 0, 1, -1, out, stats, nullptr, nullptr);
}

extern "C" int besst_bam_ingest_part(besst_ctx* ctx, const char* path, int64_t head_records, int32_t flags, int32_t part, int32_t n_parts,
                                     int64_t start_voffset, besst_records* out, besst_bam_ingest_stats* stats, int64_t* first_voffset,
                                     int64_t* landing_voffset) {
    if (!ctx) return BESST_E_INVALID;
    if (!path || !out) { ctx->err = "besst_bam_ingest: null argument"; return BESST_E_INVALID; }
    if (n_parts < 1 || part < 0 || part >= n_parts) { ctx->err = "besst_bam_ingest_part: part out of range"; return BESST_E_INVALID; }
    BESST_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const double t_start = now_s();
    if (!ctx->ingest) ctx->ingest = new BesstBamIngest();
    BesstBamIngest* S = ctx->ingest;
    for (int i = 0; i < 2; ++i) {
        if (!S->ev_h2d[i]) BESST_CUDA_TRY(ctx, cudaEventCreateWithFlags(&S->ev_h2d[i], cudaEventDisableTiming));
        if (!S->ev_inflated[i]) BESST_CUDA_TRY(ctx, cudaEventCreateWithFlags(&S->ev_inflated[i], cudaEventDisableTiming));
        if (!S->ev_scan[i]) BESST_CUDA_TRY(ctx, cudaEventCreateWithFlags(&S->ev_scan[i], cudaEventDisableTiming));
        S->h2d_pending[i] = false;
        S->inflated_pending[i] = false;
    }
    S->n = 0;
    S->n_head = 0;
    if (head_records < 0) head_records = 0;
    BESST_CUDA_TRY(ctx, S->err_d.ensure(8 * sizeof(int)));
    BESST_CUDA_TRY(ctx, S->err_h.ensure(8 * sizeof(int)));
    BESST_CUDA_TRY(ctx, S->head_rlen.ensure((size_t)(head_records + 1) * 4));
    BESST_CUDA_TRY(ctx, S->head_alen.ensure((size_t)(head_records + 1) * 4));
    if (!S->crc_ready) {
        bgzf::CrcTables t;
        bgzf::crc_make_tables(&t);
        BESST_CUDA_TRY(ctx, S->crc_d.ensure(sizeof(t)));
        BESST_CUDA_TRY(ctx, cudaMemcpy(S->crc_d.p, &t, sizeof(t), cudaMemcpyHostToDevice));
        S->crc_ready = true;
    }

    DevBackend B;
    B.ctx = ctx;
    B.S = S;
    B.flags = flags;
    B.fd = open(path, O_RDONLY);
    if (B.fd < 0) { ctx->err = std::string("besst_bam_ingest: cannot open ") + path; return BESST_E_INVALID; }
    struct stat st;
    if (fstat(B.fd, &st) != 0 || st.st_size <= 0) { close(B.fd); ctx->err = std::string("besst_bam_ingest: cannot stat / empty file: ") + path; return BESST_E_INVALID; }
    B.fsize = (int64_t)st.st_size;
    bamingest::Options opt;
    opt.head_records = head_records;
    opt.check_crc = !(flags & BESST_BAM_NO_CRC);
    opt.part = part;
    opt.n_parts = n_parts;
    opt.start_voffset = start_voffset;
    if (const char* e = getenv("BESST_BAM_TAIL")) { const long long v = atoll(e); if (v >= 0) opt.tail_bytes = v; }
    if (const char* e = getenv("BESST_BAM_WINDOW")) { const long long v = atoll(e); if (v >= 1024) opt.window_bytes = v; }
    if (const char* e = getenv("BESST_BAM_MAX_INFLATED")) { const long long v = atoll(e); if (v >= 65536) opt.max_inflated = v; }
    if (const char* e = getenv("BESST_BAM_CARRY")) { const long long v = atoll(e); if (v >= 64) opt.carry_max = (v + 3) / 4 * 4; }
    B.read_threads = std::min(12, std::max(4, (int)std::thread::hardware_concurrency()));
    if (const char* e = getenv("BESST_BAM_READ_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 64) B.read_threads = v; }
    if (const char* e = getenv("BESST_BAM_CHUNK")) { const long long v = atoll(e); if (v >= 4096) B.chunk_bytes = v; }
    opt.window_bytes = std::min<int64_t>(opt.window_bytes, std::max<int64_t>(B.fsize, 1024));

    bamingest::Result res;
    std::string why;
    int rc;
    for (;;) {
        static const int init_err[8] = {0, 0x7fffffff, 0, 0, 0, 0, 0, 0};
        if (cudaMemcpyAsync(S->err_d.p, init_err, sizeof(init_err), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { why = "H2D status"; rc = bamingest::RC_ERROR; break; }
        B.opt = opt;
        rc = bamingest::run(B, opt, &res, &why);
        if (rc == bamingest::RC_WINDOW_TOO_SMALL) {   // the header did not fit into the first window
            cudaStreamSynchronize(ctx->stream);
            cudaStreamSynchronize(ctx->copy_stream);
            opt.window_bytes *= 4;
            opt.max_inflated = std::max(opt.max_inflated, std::min<int64_t>(opt.window_bytes * 8, 3ll << 30));
            S->h2d_pending[0] = S->h2d_pending[1] = S->inflated_pending[0] = S->inflated_pending[1] = false;
            continue;
        }
        break;
    }
    close(B.fd);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    float ms[3] = {0, 0, 0};
    for (auto& e : B.timers) {
        float t = 0;
        if (cudaEventElapsedTime(&t, e.a, e.b) == cudaSuccess) ms[e.what] += t;
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    if (rc != bamingest::RC_OK) {
        ctx->err = std::string("besst_bam_ingest: ") + why + " (" + path + ")";
        cudaGetLastError();
        return BESST_E_INVALID;
    }
    S->n = res.n_records;
    S->n_head = res.n_head;
    S->ref_names.swap(res.ref_names);
    S->ref_lengths.swap(res.ref_lengths);
    besst_bam_ingest_stats& s = S->stats;
    memset(&s, 0, sizeof(s));
    s.compressed_bytes = res.stats.compressed_bytes;
    s.uncompressed_bytes = res.stats.uncompressed_bytes;
    s.blocks = res.stats.blocks;
    s.records = res.stats.records;
    s.windows = res.stats.windows;
    s.rescans = res.stats.rescans;
    s.ms_inflate = ms[0];
    s.ms_scan = ms[1];
    s.ms_decode = ms[2];
    s.seconds_read = B.t_read;
    s.seconds_total = now_s() - t_start;
    s.crc_checked = opt.check_crc ? 1 : 0;
    if (stats) *stats = s;
    if (first_voffset) *first_voffset = res.first_voffset;
    if (landing_voffset) *landing_voffset = res.landing_voffset;
    memset(out, 0, sizeof(*out));
    out->n = S->n;
    out->tid = S->col_i32[0].as<int32_t>(); out->mtid = S->col_i32[1].as<int32_t>(); out->pos = S->col_i32[2].as<int32_t>();
    out->mpos = S->col_i32[3].as<int32_t>(); out->tlen = S->col_i32[4].as<int32_t>(); out->qlen = S->col_i32[5].as<int32_t>();
    out->flag = S->col_flag.as<uint16_t>(); out->mapq = S->col_mapq.as<uint8_t>();
    out->packed = S->unpackable ? nullptr : S->col_packed.as<uint32_t>();
    out->on_device = 1;
    return BESST_OK;
}

extern "C" int64_t besst_bam_ingest_n_refs(besst_ctx* ctx) { return ctx && ctx->ingest ? (int64_t)ctx->ingest->ref_names.size() : -1; }
extern "C" const char* besst_bam_ingest_ref_name(besst_ctx* ctx, int64_t i) {
    return ctx && ctx->ingest && i >= 0 && i < (int64_t)ctx->ingest->ref_names.size() ? ctx->ingest->ref_names[(size_t)i].c_str() : nullptr;
}
extern "C" int64_t besst_bam_ingest_ref_length(besst_ctx* ctx, int64_t i) {
    return ctx && ctx->ingest && i >= 0 && i < (int64_t)ctx->ingest->ref_lengths.size() ? ctx->ingest->ref_lengths[(size_t)i] : -1;
}
extern "C" int64_t besst_bam_ingest_head(besst_ctx* ctx, int32_t* rlen, int32_t* alen, int64_t cap) {
    if (!ctx || !ctx->ingest) return BESST_E_STATE;
    BesstBamIngest* S = ctx->ingest;
    const int64_t n = std::min<int64_t>(S->n_head, cap);
    if (n > 0) {
        if (rlen && cudaMemcpyAsync(rlen, S->head_rlen.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return BESST_E_CUDA;
        if (alen && cudaMemcpyAsync(alen, S->head_alen.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return BESST_E_CUDA;
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return BESST_E_CUDA;
    }
    return n;
}
extern "C" int besst_device_read(besst_ctx* ctx, const void* device_ptr, void* host_ptr, int64_t bytes) {
    if (!ctx || (bytes > 0 && (!device_ptr || !host_ptr))) return BESST_E_INVALID;
    if (bytes <= 0) return BESST_OK;
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(host_ptr, device_ptr, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BESST_OK;
}
