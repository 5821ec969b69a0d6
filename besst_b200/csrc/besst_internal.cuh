// Internal declarations shared by the .cu translation units of libbesst_b200.so.
// Nothing here crosses the C ABI (include/besst_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/besst_b200.h"

#define BESST_CUDA_TRY(ctx, expr)                                                           \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                \
            return BESST_E_CUDA;                                                            \
        }                                                                                   \
    } while (0)

// grow-only device buffer
struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// grow-only pinned host buffer (results handed out as views, see besst_graph_view)
struct HBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct DeviceRecords {
    int64_t n;
    const int32_t *tid, *mtid, *pos, *mpos, *tlen, *qlen;
    const uint16_t* flag;
    const uint8_t* mapq;
    const uint32_t* packed = nullptr;   // flag | mapq << 12 | qlen << 20 (replaces the three columns in the graph build)
};

// scalar constants of the per-edge scoring math, evaluated once on the host
// with the same libm calls the oracle makes (pow), so both sides start from
// identical fp64 values
struct ScoreConsts {
    double mean, sd, r;
    double s2;        // 2**0.5 * sd
    double v2;        // 2 * sd**2
    double k;         // sd / (2*pi)**0.5
    double gb_den;    // (2*pi)**0.5 * sd
    double sd2, sd4, mean2;
    double d_upper0, d_lower0;  // bisection bracket of GapEstimator
    int erf_variant;
};

struct BesstBamIngest;   // besst_bamdev.cu: columns and window buffers of besst_bam_ingest

struct besst_ctx {
    int device = 0;
    BesstBamIngest* ingest = nullptr;
    int sm_count = 148;
    cudaStream_t stream = nullptr;      // the stream all work is ordered on
    cudaStream_t own_stream = nullptr;  // created by besst_create
    bool use_caller_stream = false;
    std::string err;

    // contig table (the current one); besst_contigs_select parks it in `tables` and swaps another one in
    DBuf rows, rows_packed, scaf_len;
    int64_t n_contigs = 0, n_scaffolds = 0, n_large = 0;
    struct TableSlot { DBuf rows, rows_packed, scaf_len; int64_t n_contigs = 0, n_scaffolds = 0, n_large = 0; };
    TableSlot tables[BESST_MAX_TABLES];
    int cur_table = 0;

    // staged records (host-pointer calls); the copies run on copy_stream, slice by slice, overlapped
    // with the record kernel on `stream`
    DBuf rec_i32[6], rec_flag, rec_mapq, rec_packed;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_fork = nullptr;
    std::vector<cudaEvent_t> slice_events;
    int64_t slice_records = 16ll << 20;   // records per H2D slice (multiple of 128)

    // pinned host copies of the result (besst_graph_view)
    HBuf h_out[17];

    // link extraction
    DBuf tuples, scratch_tuples, block_tile0, tile_aggs, fishy_keys, aligned, counters, tile_state, part_state, misc;
    int64_t fishy_cap = 0;
    int64_t n_tuples = 0, n_fishy_keys = 0;
    int64_t n_rec_tiles = 0;     // 128-record tiles of the last extraction (scratch_tuples / tile_state layout)
    bool have_links = false;
    bool tuples_valid = false;   // `tuples` holds the BAM-ordered array (else the links still sit in scratch_tuples)

    // sort
    DBuf key_a, key_b, idx_a, idx_b, hist, sort_state;

    // CSR / per-edge results
    DBuf fishy_sorted, fishy_tmp, heads, block_sums;
    DBuf e_u, e_v, e_nr, e_obs, e_obs_sq, e_first, e_row_ptr, e_gap, e_score, e_ks, e_sd_obs, e_sd_model, e_fishy,
        e_flags, l_obs_u, l_obs_v, e_sum_u, e_max_v, ll_off, ks_key[6];
    // run-merge bucket: block-grouped observations and the run descriptors
    DBuf grouped, run_key[2], run_val[2], run_start, run_cnt, run_first, run_off, run_src, run_len, edge_run_ptr, run_state;
    int64_t n_edges = 0, n_links = 0, n_fishy_pairs = 0, n_ll_links = 0;
    int64_t n_runs = 0;        // run descriptors of the last besst_links_group
    int run_block_bits = 0;
    bool have_runs = false;
    bool attr_group_done = false;
    // pinned host words for the size read-backs between kernels (pageable targets go through a bounce buffer)
    uint64_t* h_scalars = nullptr;
    uint64_t* host_scalars() {
        if (!h_scalars && cudaHostAlloc(reinterpret_cast<void**>(&h_scalars), 64 * sizeof(uint64_t), cudaHostAllocDefault) != cudaSuccess)
            h_scalars = nullptr;
        return h_scalars;
    }
    bool have_graph = false;
    besst_lib_params last_params;
    besst_lib_params extract_params;   // of the last besst_links_extract

    // per-kernel profiling (optional)
    bool prof = false;
    bool prof_accumulate = false;   // keep the launches of several builds until besst_kernel_profile reads them
    struct ProfEntry { int id; cudaEvent_t a, b; };
    std::vector<ProfEntry> prof_pool;
    size_t prof_used = 0;

    // timing
    cudaEvent_t ev[BESST_N_STAGES + 1];
    bool ev_valid = false;
    int n_stage_marks = 0;
    int64_t launches = 0;
    int sweep_kernel_id = BESST_K_RADIX_SWEEP;  // profiling id of the next radix sort's digit passes
};

// destination rank of an edge in the multi-GPU exchange: hash(u, v) mod world (tuple-level and run-level paths)
#ifdef __CUDACC__
__device__ __forceinline__ unsigned int besst_edge_dest(unsigned int u, unsigned int v, int world) {
    unsigned long long x = ((unsigned long long)u << 32) | v;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (unsigned int)(x % (unsigned long long)world);
}
#endif

void besst_bamdev_release(besst_ctx* ctx);

// ---- launchers (one per translation unit) ----------------------------------
// links: records -> accepted link tuples (BAM order), coverage, fishy keys, counters
int besst_launch_extract(besst_ctx* ctx, const besst_lib_params& p, const DeviceRecords& rec);
// the same in phases, for callers that feed the record range in slices (H2D overlap)
int besst_extract_begin(besst_ctx* ctx, const besst_lib_params& p, int64_t n);
int besst_extract_slice(besst_ctx* ctx, const besst_lib_params& p, const DeviceRecords& rec, int64_t r0, int64_t r1);
int besst_extract_finish(besst_ctx* ctx, const besst_lib_params& p, int64_t n, bool* overflow);
int besst_ensure_tuples(besst_ctx* ctx);

// sort + CSR: tuples -> sorted (key, idx) -> edges
int besst_launch_graph(besst_ctx* ctx, const besst_lib_params& p, const besst_link_tuple* d_tuples, int64_t n_tuples,
                       const uint64_t* d_fishy, int64_t n_fishy);

// the graph build from run descriptors prepared by the caller (multi-GPU import of exchanged runs)
struct BesstRunInput {
    const int2* grouped;   // (obs_u, obs_v) per link; run r covers [run_start[r], run_start[r] + run_cnt[r])
    int64_t n_runs;        // descriptors in ctx->run_key[0] / run_val[0] / run_start / run_cnt / run_first
    int low_bits;          // bits of the sort key below the edge key (source rank, block)
    bool packed16;         // one u32 per link (obs_u | obs_v << 16) instead of an int2
};
// bytes per link of the exchanged observations under these library parameters: every accepted observation is
// below ins_size_threshold (CreateGraph.py:840), so two of them share a 32-bit word when it is <= 65535
static inline int besst_obs_bytes(const besst_lib_params& p) {
    return (p.ins_size_threshold > 0 && p.ins_size_threshold <= 65535.0) ? 4 : 8;
}
int besst_launch_graph_from_runs(besst_ctx* ctx, const besst_lib_params& p, int64_t n_links, const BesstRunInput& runs,
                                 const uint64_t* d_fishy, int64_t n_fishy);
int besst_launch_runs_route(besst_ctx* ctx, int world, int64_t* link_counts, int64_t* run_counts);
int besst_launch_runs_route_async(besst_ctx* ctx, int world, int block_bits, int64_t run_cap);
int besst_launch_partition_fishy_async(besst_ctx* ctx, int world, uint64_t* out_fishy, const uint64_t** totals_device);
int besst_launch_runs_pack(besst_ctx* ctx, int world, int32_t* out_obs, besst_run_desc* out_desc, int32_t* const* obs_ptrs,
                           besst_run_desc* const* desc_ptrs);
int besst_launch_runs_import(besst_ctx* ctx, const besst_run_desc* desc, int64_t n_runs, int world, int block_bits,
                             const int64_t* src_run_counts, const int64_t* src_link_counts, const int64_t* src_first_base,
                             int* low_bits);
// d_tuples == nullptr: group the last extraction straight from its tile-local scratch runs
int besst_group_tuples(besst_ctx* ctx, const besst_link_tuple* d_tuples, int64_t n, int bv, int block_bits, int64_t* n_runs,
                       int* overflow);

// radix sort of 64-bit keys with 32-bit payload (onesweep); returns which buffer holds the result
int besst_radix_sort_pairs(besst_ctx* ctx, uint64_t* keys_a, uint64_t* keys_b, uint32_t* val_a, uint32_t* val_b,
                           int64_t n, int key_bits, int* result_in_b);
int besst_radix_sort_keys(besst_ctx* ctx, uint64_t* keys_a, uint64_t* keys_b, int64_t n, int key_bits,
                          int* result_in_b);

// per-edge statistics, KS, GapEst, score
int besst_launch_edge_stats(besst_ctx* ctx, const besst_lib_params& p, const besst_link_tuple* d_tuples,
                            const uint64_t* d_sorted_keys, const uint32_t* d_sorted_idx, int64_t n_links, int key_shift);
int besst_launch_gapest(besst_ctx* ctx, const besst_lib_params& p, const double* d_mean_obs, const double* d_len1,
                        const double* d_len2, int64_t n, int32_t* d_gap, double* d_sd);
int besst_launch_trsk_sd(besst_ctx* ctx, const besst_lib_params& p, const double* d_gap, const double* d_len1,
                         const double* d_len2, int64_t n, double* d_sd);
int besst_launch_func_of_d(besst_ctx* ctx, const besst_lib_params& p, const double* d_d, const double* d_len1,
                           const double* d_len2, int64_t n, double* d_out);
int besst_launch_gapest_lognormal(besst_ctx* ctx, double mu, double sigma, double r, const int32_t* d_samples, const int64_t* d_row_ptr,
                                  const double* d_len1, const double* d_len2, int64_t n, int32_t* d_gap);
int besst_launch_partition(besst_ctx* ctx, int world, besst_link_tuple* out_tuples, uint32_t* out_ordinals,
                           uint64_t* out_fishy, int64_t* tuple_counts, int64_t* fishy_counts);

ScoreConsts besst_score_consts(const besst_lib_params& p);

int besst_launch_libmetrics(besst_ctx* ctx, const besst_lib_params& p, const DeviceRecords& rec,
                            const int64_t* ref_lengths, int64_t n_refs, int32_t want_isize, besst_libmetrics_out* out,
                            double* adjusted_distribution, int64_t cap);

static inline void besst_mark(besst_ctx* ctx) {
    if (ctx->n_stage_marks <= BESST_N_STAGES) cudaEventRecord(ctx->ev[ctx->n_stage_marks++], ctx->stream);
}

// RAII event pair around one kernel launch (no-op unless profiling is enabled)
struct KTimer {
    besst_ctx* c;
    size_t slot;
    bool on;
    KTimer(besst_ctx* ctx, int id) : c(ctx), slot(0), on(ctx->prof) {
        c->launches++;
        if (!on) return;
        if (c->prof_used == c->prof_pool.size()) {
            besst_ctx::ProfEntry e;
            e.id = id;
            cudaEventCreate(&e.a);
            cudaEventCreate(&e.b);
            c->prof_pool.push_back(e);
        }
        slot = c->prof_used++;
        c->prof_pool[slot].id = id;
        cudaEventRecord(c->prof_pool[slot].a, c->stream);
    }
    ~KTimer() {
        if (on) cudaEventRecord(c->prof_pool[slot].b, c->stream);
    }
};
