// C ABI of libbesst_b200.so (include/besst_b200.h).  Host-side glue only: device
// buffers, H2D/D2H marshalling, stage ordering on one CUDA stream, CUDA-event
// timing.  There is no CPU implementation behind these entry points: without a
// CUDA device besst_create fails.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "besst_internal.cuh"

static thread_local std::string g_create_error;

extern "C" int besst_abi_version(void) { return BESST_ABI_VERSION; }

extern "C" besst_ctx* besst_create(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return nullptr;
    }
    if (device >= 0) {
        e = cudaSetDevice(device);
        if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return nullptr; }
    } else {
        cudaGetDevice(&device);
    }
    besst_ctx* ctx = new besst_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e); delete ctx; return nullptr; }
    ctx->stream = ctx->own_stream;
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (const char* e = getenv("BESST_SLICE_RECORDS")) {   // tuning knob: records per H2D slice
        const long long v = atoll(e);
        if (v >= 128) ctx->slice_records = (v / 128) * 128;
    }
    for (int i = 0; i <= BESST_N_STAGES; ++i) cudaEventCreate(&ctx->ev[i]);
    return ctx;
}

extern "C" void besst_destroy(besst_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    besst_bamdev_release(ctx);
    DBuf* bufs[] = {&ctx->rows, &ctx->rows_packed, &ctx->scratch_tuples, &ctx->block_tile0, &ctx->tile_aggs, &ctx->part_state, &ctx->scaf_len, &ctx->rec_flag, &ctx->rec_mapq, &ctx->rec_packed, &ctx->tuples, &ctx->fishy_keys, &ctx->aligned,
                    &ctx->tile_state, &ctx->misc, &ctx->key_a, &ctx->key_b, &ctx->idx_a, &ctx->idx_b, &ctx->hist,
                    &ctx->sort_state, &ctx->fishy_sorted, &ctx->fishy_tmp, &ctx->heads, &ctx->block_sums, &ctx->e_u, &ctx->e_v,
                    &ctx->e_nr, &ctx->e_obs, &ctx->e_obs_sq, &ctx->e_first, &ctx->e_row_ptr, &ctx->e_gap, &ctx->e_score,
                    &ctx->e_ks, &ctx->e_sd_obs, &ctx->e_sd_model, &ctx->e_fishy, &ctx->e_flags, &ctx->l_obs_u, &ctx->l_obs_v,
                    &ctx->e_sum_u, &ctx->e_max_v, &ctx->ll_off, &ctx->ks_key[0], &ctx->ks_key[1], &ctx->ks_key[2], &ctx->ks_key[3], &ctx->ks_key[4], &ctx->ks_key[5],
                    &ctx->grouped, &ctx->run_key[0], &ctx->run_key[1], &ctx->run_val[0], &ctx->run_val[1], &ctx->run_start, &ctx->run_cnt,
                    &ctx->run_first, &ctx->run_off, &ctx->run_src, &ctx->run_len, &ctx->edge_run_ptr, &ctx->run_state};
    for (DBuf* b : bufs) b->release();
    for (auto& t : ctx->tables) { t.rows.release(); t.rows_packed.release(); t.scaf_len.release(); }
    for (DBuf& b : ctx->rec_i32) b.release();
    for (HBuf& b : ctx->h_out) b.release();
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    for (cudaEvent_t e : ctx->slice_events) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i <= BESST_N_STAGES; ++i) cudaEventDestroy(ctx->ev[i]);
    for (auto& e : ctx->prof_pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" const char* besst_last_error(besst_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int besst_set_contigs(besst_ctx* ctx, const besst_contig_row* rows, int64_t n_contigs, int64_t n_scaffolds,
                                 int64_t n_large_scaffolds) {
    if (!ctx) return BESST_E_INVALID;
    if (n_contigs < 0 || n_scaffolds < 0 || n_large_scaffolds < 0 || n_large_scaffolds > n_scaffolds || (n_contigs > 0 && !rows) ||
        n_scaffolds >= (1ll << 29) || n_contigs >= (1ll << 31)) {
        ctx->err = "besst_set_contigs: bad sizes";
        return BESST_E_INVALID;
    }
    std::vector<int32_t> slen((size_t)(n_scaffolds > 0 ? n_scaffolds : 1), 0);
    // packed 16-byte rows for the record kernel: one 128-bit gather per read end
    //   x = scaffold << 3 | direction << 2 | state,  y = position,  z = length,  w = scaf_length
    std::vector<int32_t> packed(4 * (size_t)(n_contigs > 0 ? n_contigs : 1), 0);
    for (int64_t c = 0; c < n_contigs; ++c) {
        if (rows[c].state == BESST_CTG_ABSENT) continue;
        if (rows[c].state != BESST_CTG_LARGE && rows[c].state != BESST_CTG_SMALL) { ctx->err = "besst_set_contigs: bad contig state"; return BESST_E_INVALID; }
        if (rows[c].scaffold < 0 || rows[c].scaffold >= n_scaffolds) { ctx->err = "besst_set_contigs: scaffold index out of range"; return BESST_E_INVALID; }
        slen[(size_t)rows[c].scaffold] = rows[c].scaf_length;
        packed[4 * (size_t)c + 0] = (int32_t)(((uint32_t)rows[c].scaffold << 3) | (rows[c].direction ? 4u : 0u) | (uint32_t)rows[c].state);
        packed[4 * (size_t)c + 1] = rows[c].position;
        packed[4 * (size_t)c + 2] = rows[c].length;
        packed[4 * (size_t)c + 3] = rows[c].scaf_length;
    }
    cudaSetDevice(ctx->device);
    BESST_CUDA_TRY(ctx, ctx->rows.ensure(sizeof(besst_contig_row) * (size_t)(n_contigs > 0 ? n_contigs : 1)));
    BESST_CUDA_TRY(ctx, ctx->scaf_len.ensure(4 * slen.size()));
    BESST_CUDA_TRY(ctx, ctx->rows_packed.ensure(4 * packed.size()));
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->rows_packed.p, packed.data(), 4 * packed.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (n_contigs > 0)
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->rows.p, rows, sizeof(besst_contig_row) * (size_t)n_contigs, cudaMemcpyHostToDevice, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scaf_len.p, slen.data(), 4 * slen.size(), cudaMemcpyHostToDevice, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_contigs = n_contigs; ctx->n_scaffolds = n_scaffolds; ctx->n_large = n_large_scaffolds;
    ctx->have_links = ctx->have_graph = false;
    return BESST_OK;
}

extern "C" int besst_contigs_select(besst_ctx* ctx, int32_t slot) {
    if (!ctx) return BESST_E_INVALID;
    if (slot < 0 || slot >= BESST_MAX_TABLES) { ctx->err = "besst_contigs_select: slot out of range"; return BESST_E_INVALID; }
    if (slot == ctx->cur_table) return BESST_OK;
    besst_ctx::TableSlot& out = ctx->tables[ctx->cur_table];
    besst_ctx::TableSlot& in = ctx->tables[slot];
    std::swap(out.rows, ctx->rows); std::swap(out.rows_packed, ctx->rows_packed); std::swap(out.scaf_len, ctx->scaf_len);
    out.n_contigs = ctx->n_contigs; out.n_scaffolds = ctx->n_scaffolds; out.n_large = ctx->n_large;
    std::swap(in.rows, ctx->rows); std::swap(in.rows_packed, ctx->rows_packed); std::swap(in.scaf_len, ctx->scaf_len);
    ctx->n_contigs = in.n_contigs; ctx->n_scaffolds = in.n_scaffolds; ctx->n_large = in.n_large;
    ctx->cur_table = slot;
    ctx->have_links = ctx->have_graph = ctx->have_runs = false;
    return BESST_OK;
}

// host -> device staging of a record batch (no-op for device-resident batches)
static int stage_records(besst_ctx* ctx, const besst_records* r, DeviceRecords* d, bool need_tlen, bool need_graph_cols) {
    if (!r || r->n < 0) { ctx->err = "records: null or negative n"; return BESST_E_INVALID; }
    d->n = r->n;
    const bool packed = need_graph_cols && r->packed != nullptr;   // the graph build takes flag / mapq / qlen from the packed column
    if (r->n > 0 && (!r->tid || !r->mtid || (!packed && (!r->flag || !r->mapq)) || (need_tlen && !r->tlen) ||
                     (need_graph_cols && (!r->pos || !r->mpos || (!packed && !r->qlen))))) {
        ctx->err = "records: missing column";
        return BESST_E_INVALID;
    }
    d->packed = nullptr;
    if (r->on_device) {
        d->tid = r->tid; d->mtid = r->mtid; d->pos = r->pos; d->mpos = r->mpos; d->tlen = r->tlen; d->qlen = r->qlen;
        d->flag = r->flag; d->mapq = r->mapq;
        if (packed) d->packed = r->packed;
        return BESST_OK;
    }
    const size_t n = (size_t)r->n, nz = n ? n : 1;
    const int32_t* src[6] = {r->tid, r->mtid, r->pos, r->mpos, r->tlen, packed ? nullptr : r->qlen};
    const int32_t** dst[6] = {&d->tid, &d->mtid, &d->pos, &d->mpos, &d->tlen, &d->qlen};
    for (int k = 0; k < 6; ++k) {
        *dst[k] = nullptr;
        if (!src[k]) continue;
        BESST_CUDA_TRY(ctx, ctx->rec_i32[k].ensure(4 * nz));
        if (n) BESST_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->rec_i32[k].p, src[k], 4 * n, cudaMemcpyHostToDevice, ctx->stream));
        *dst[k] = ctx->rec_i32[k].as<int32_t>();
    }
    if (packed) {
        BESST_CUDA_TRY(ctx, ctx->rec_packed.ensure(4 * nz));
        if (n) BESST_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->rec_packed.p, r->packed, 4 * n, cudaMemcpyHostToDevice, ctx->stream));
        d->packed = ctx->rec_packed.as<uint32_t>();
        d->flag = nullptr; d->mapq = nullptr;
        return BESST_OK;
    }
    BESST_CUDA_TRY(ctx, ctx->rec_flag.ensure(2 * nz));
    BESST_CUDA_TRY(ctx, ctx->rec_mapq.ensure(nz));
    if (n) {
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->rec_flag.p, r->flag, 2 * n, cudaMemcpyHostToDevice, ctx->stream));
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->rec_mapq.p, r->mapq, n, cudaMemcpyHostToDevice, ctx->stream));
    }
    d->flag = ctx->rec_flag.as<uint16_t>();
    d->mapq = ctx->rec_mapq.as<uint8_t>();
    return BESST_OK;
}

// Host-buffer graph build: the seven record columns are copied slice by slice on copy_stream while
// the record kernel works on the slices that have arrived (an event per slice orders the two
// streams), so the call costs max(PCIe, K1) instead of their sum.
static int extract_host_pipelined(besst_ctx* ctx, const besst_lib_params* params, const besst_records* r, DeviceRecords* d) {
    const bool packed = r->packed != nullptr;
    if (r->n < 0 || !r->tid || !r->mtid || !r->pos || !r->mpos || (!packed && (!r->flag || !r->mapq || !r->qlen))) {
        ctx->err = "records: missing column";
        return BESST_E_INVALID;
    }
    const size_t n = (size_t)r->n;
    const int32_t* src[6] = {r->tid, r->mtid, r->pos, r->mpos, nullptr, packed ? nullptr : r->qlen};
    const int32_t** dst[6] = {&d->tid, &d->mtid, &d->pos, &d->mpos, &d->tlen, &d->qlen};
    for (int k = 0; k < 6; ++k) {
        *dst[k] = nullptr;
        if (!src[k]) continue;
        BESST_CUDA_TRY(ctx, ctx->rec_i32[k].ensure(4 * n));
        *dst[k] = ctx->rec_i32[k].as<int32_t>();
    }
    d->packed = nullptr; d->flag = nullptr; d->mapq = nullptr;
    if (packed) {
        BESST_CUDA_TRY(ctx, ctx->rec_packed.ensure(4 * n));
        d->packed = ctx->rec_packed.as<uint32_t>();
    } else {
        BESST_CUDA_TRY(ctx, ctx->rec_flag.ensure(2 * n));
        BESST_CUDA_TRY(ctx, ctx->rec_mapq.ensure(n));
        d->flag = ctx->rec_flag.as<uint16_t>();
        d->mapq = ctx->rec_mapq.as<uint8_t>();
    }
    d->n = r->n;
    const int64_t S = ctx->slice_records;
    const int64_t n_slices = (r->n + S - 1) / S;
    while ((int64_t)ctx->slice_events.size() < n_slices) {
        cudaEvent_t e;
        BESST_CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->slice_events.push_back(e);
    }
    int rc = besst_extract_begin(ctx, *params, r->n);
    if (rc) return rc;
    // the staging buffers may still be read by earlier work on `stream`
    BESST_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
    for (int64_t s = 0; s < n_slices; ++s) {
        const size_t r0 = (size_t)(s * S), r1 = (size_t)((s + 1) * S < r->n ? (s + 1) * S : r->n), m = r1 - r0;
        for (int k = 0; k < 6; ++k)
            if (src[k])
                BESST_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->rec_i32[k].as<int32_t>() + r0, src[k] + r0, 4 * m, cudaMemcpyHostToDevice, ctx->copy_stream));
        if (packed) {
            BESST_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->rec_packed.as<uint32_t>() + r0, r->packed + r0, 4 * m, cudaMemcpyHostToDevice, ctx->copy_stream));
        } else {
            BESST_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->rec_flag.as<uint16_t>() + r0, r->flag + r0, 2 * m, cudaMemcpyHostToDevice, ctx->copy_stream));
            BESST_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->rec_mapq.as<uint8_t>() + r0, r->mapq + r0, m, cudaMemcpyHostToDevice, ctx->copy_stream));
        }
        BESST_CUDA_TRY(ctx, cudaEventRecord(ctx->slice_events[(size_t)s], ctx->copy_stream));
        BESST_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->slice_events[(size_t)s], 0));
        rc = besst_extract_slice(ctx, *params, *d, (int64_t)r0, (int64_t)r1);
        if (rc) return rc;
    }
    bool overflow = false;
    rc = besst_extract_finish(ctx, *params, r->n, &overflow);
    if (rc) return rc;
    if (overflow) return besst_launch_extract(ctx, *params, *d);   // the records are resident now
    return BESST_OK;
}

static int check_params(besst_ctx* ctx, const besst_lib_params* p) {
    if (!p) { ctx->err = "params: null"; return BESST_E_INVALID; }
    if (p->orientation != BESST_ORIENT_FR && p->orientation != BESST_ORIENT_RF) { ctx->err = "params: orientation must be fr or rf"; return BESST_E_INVALID; }
    return BESST_OK;
}

extern "C" int besst_links_extract(besst_ctx* ctx, const besst_lib_params* params, const besst_records* records, int64_t* n_tuples) {
    if (!ctx) return BESST_E_INVALID;
    int rc = check_params(ctx, params);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    ctx->have_links = ctx->have_graph = ctx->have_runs = false;
    ctx->extract_params = *params;
    ctx->n_stage_marks = 0;
    if (!ctx->prof_accumulate) ctx->prof_used = 0;
    besst_mark(ctx);
    DeviceRecords d;
    if (records && !records->on_device && records->n > ctx->slice_records) {
        rc = extract_host_pipelined(ctx, params, records, &d);
    } else {
        rc = stage_records(ctx, records, &d, false, true);
        if (rc) return rc;
        rc = besst_launch_extract(ctx, *params, d);
    }
    if (rc) return rc;
    besst_mark(ctx);
    if (n_tuples) *n_tuples = ctx->n_tuples;
    return BESST_OK;
}

extern "C" int besst_links_tuples_device(besst_ctx* ctx, const besst_link_tuple** tuples, int64_t* n_tuples) {
    if (!ctx || !ctx->have_links) { if (ctx) ctx->err = "no extracted links"; return BESST_E_STATE; }
    cudaSetDevice(ctx->device);
    { const int rc = besst_ensure_tuples(ctx); if (rc) return rc; }
    *tuples = ctx->tuples.as<besst_link_tuple>();
    *n_tuples = ctx->n_tuples;
    return BESST_OK;
}

extern "C" int besst_links_fishy_device(besst_ctx* ctx, const uint64_t** keys, int64_t* n_keys) {
    if (!ctx || !ctx->have_links) { if (ctx) ctx->err = "no extracted links"; return BESST_E_STATE; }
    *keys = ctx->fishy_keys.as<uint64_t>();
    *n_keys = ctx->n_fishy_keys;
    return BESST_OK;
}

extern "C" int besst_links_partials(besst_ctx* ctx, int64_t* aligned_len_host, int64_t* counters_host) {
    if (!ctx || !ctx->have_links) { if (ctx) ctx->err = "no extracted links"; return BESST_E_STATE; }
    cudaSetDevice(ctx->device);
    if (aligned_len_host && ctx->n_contigs > 0)
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(aligned_len_host, ctx->aligned.p, 8 * (size_t)ctx->n_contigs, cudaMemcpyDeviceToHost, ctx->stream));
    if (counters_host)
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(counters_host, ctx->counters.p, 8 * BESST_N_COUNTERS, cudaMemcpyDeviceToHost, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BESST_OK;
}

extern "C" int besst_links_partials_device(besst_ctx* ctx, int64_t** aligned_len_device, int64_t** counters_device) {
    if (!ctx || !ctx->have_links) { if (ctx) ctx->err = "no extracted links"; return BESST_E_STATE; }
    if (aligned_len_device) *aligned_len_device = ctx->aligned.as<int64_t>();
    if (counters_device) *counters_device = ctx->counters.as<int64_t>();
    return BESST_OK;
}

extern "C" int besst_links_fetch(besst_ctx* ctx, besst_link_tuple* tuples_host, uint64_t* fishy_keys_host) {
    if (!ctx || !ctx->have_links) { if (ctx) ctx->err = "no extracted links"; return BESST_E_STATE; }
    cudaSetDevice(ctx->device);
    if (tuples_host) { const int rc = besst_ensure_tuples(ctx); if (rc) return rc; }
    if (tuples_host && ctx->n_tuples > 0)
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(tuples_host, ctx->tuples.p, sizeof(besst_link_tuple) * (size_t)ctx->n_tuples, cudaMemcpyDeviceToHost, ctx->stream));
    if (fishy_keys_host && ctx->n_fishy_keys > 0)
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(fishy_keys_host, ctx->fishy_keys.p, 8 * (size_t)ctx->n_fishy_keys, cudaMemcpyDeviceToHost, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BESST_OK;
}

extern "C" int besst_links_partition(besst_ctx* ctx, int32_t world, besst_link_tuple* out_tuples_device,
                                     uint32_t* out_ordinals_device, uint64_t* out_fishy_device, int64_t* tuple_counts,
                                     int64_t* fishy_counts) {
    if (!ctx || !ctx->have_links) { if (ctx) ctx->err = "no extracted links"; return BESST_E_STATE; }
    if (!tuple_counts || !fishy_counts || (ctx->n_fishy_keys > 0 && !out_fishy_device)) {
        ctx->err = "links_partition: bad arguments";
        return BESST_E_INVALID;
    }
    cudaSetDevice(ctx->device);
    return besst_launch_partition(ctx, world, out_tuples_device, out_ordinals_device, out_fishy_device, tuple_counts, fishy_counts);
}

extern "C" int besst_links_group(besst_ctx* ctx, int64_t* n_runs) {
    if (!ctx || !ctx->have_links) { if (ctx) ctx->err = "no extracted links"; return BESST_E_STATE; }
    cudaSetDevice(ctx->device);
    ctx->have_runs = false;
    int bv = 1;
    {
        const uint64_t m = (uint64_t)(2 * ctx->n_scaffolds > 0 ? 2 * ctx->n_scaffolds - 1 : 1);
        while (bv < 32 && (m >> bv)) ++bv;
    }
    const int64_t n = ctx->n_tuples;
    const int64_t blocks = (n + 2047) / 2048;
    int bb = 1;
    while (bb < 32 && ((uint64_t)(blocks > 1 ? blocks - 1 : 1) >> bb)) ++bb;
    if (2 * bv + bb > 64) return 1;
    int64_t R = 0;
    int overflow = 0;
    int rc = besst_group_tuples(ctx, nullptr, n, bv, bb, &R, &overflow);
    if (rc) return rc;
    if (overflow) return 1;
    ctx->n_runs = R;
    ctx->run_block_bits = bb;
    ctx->have_runs = true;
    if (n_runs) *n_runs = R;
    return BESST_OK;
}

// group + route counts + fishy partition queued back to back, ONE host read for all sizes
extern "C" int besst_exchange_prepare(besst_ctx* ctx, int32_t world, uint64_t* out_fishy_device, int64_t* summary,
                                      int64_t* link_counts, int64_t* run_counts, int64_t* fishy_counts) {
    if (!ctx || !ctx->have_links) { if (ctx) ctx->err = "no extracted links"; return BESST_E_STATE; }
    if (world < 1 || world > 16 || !summary || !link_counts || !run_counts || !fishy_counts || (ctx->n_fishy_keys > 0 && !out_fishy_device)) {
        ctx->err = "exchange_prepare: bad arguments";
        return BESST_E_INVALID;
    }
    cudaSetDevice(ctx->device);
    ctx->have_runs = false;
    int bv = 1;
    {
        const uint64_t m = (uint64_t)(2 * ctx->n_scaffolds > 0 ? 2 * ctx->n_scaffolds - 1 : 1);
        while (bv < 32 && (m >> bv)) ++bv;
    }
    const int64_t n = ctx->n_tuples;
    const int64_t blocks = (n + 2047) / 2048;
    int bb = 1;
    while (bb < 32 && ((uint64_t)(blocks > 1 ? blocks - 1 : 1) >> bb)) ++bb;
    const bool keys_fit = 2 * bv + bb <= 64;
    const int64_t run_cap = n / 8 > (1 << 16) ? n / 8 : (1 << 16);
    int rc;
    if (keys_fit) {
        rc = besst_group_tuples(ctx, nullptr, n, bv, bb, nullptr, nullptr);
        if (rc) return rc;
        rc = besst_launch_runs_route_async(ctx, world, bb, run_cap);
        if (rc) return rc;
    }
    const uint64_t* fishy_totals = nullptr;
    rc = besst_launch_partition_fishy_async(ctx, world, out_fishy_device, &fishy_totals);
    if (rc) return rc;
    uint64_t* const h = ctx->host_scalars();
    if (!h) { ctx->err = "pinned host scratch allocation failed"; return BESST_E_NOMEM; }
    memset(h, 0, 64 * sizeof(uint64_t));
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->counters.p, 8 * BESST_N_COUNTERS, cudaMemcpyDeviceToHost, ctx->stream));
    if (keys_fit) {
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(h + 16, ctx->run_state.p, 8, cudaMemcpyDeviceToHost, ctx->stream));                          // runs, overflow flags
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(h + 17, ctx->run_state.as<uint32_t>() + 16, 4 * 32, cudaMemcpyDeviceToHost, ctx->stream));    // route counts
    }
    if (fishy_totals) BESST_CUDA_TRY(ctx, cudaMemcpyAsync(h + 33, fishy_totals, 8 * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t* gs = reinterpret_cast<const uint32_t*>(h + 16);
    const uint32_t* route = reinterpret_cast<const uint32_t*>(h + 17);
    const int64_t R = gs[0];
    const bool ok = keys_fit && gs[1] == 0 && R <= run_cap;
    summary[0] = ok ? 1 : 0;
    summary[1] = ok ? R : 0;
    summary[2] = (int64_t)h[BESST_CNT_CALLS];
    summary[3] = (int64_t)h[BESST_CNT_LAST_OBS1]; summary[4] = (int64_t)h[BESST_CNT_LAST_OBS2];
    summary[5] = (int64_t)h[BESST_CNT_FIRST_OBS1]; summary[6] = (int64_t)h[BESST_CNT_FIRST_OBS2];
    summary[7] = n;
    for (int d = 0; d < world; ++d) {
        link_counts[d] = ok ? route[d] : 0;
        run_counts[d] = ok ? route[16 + d] : 0;
        fishy_counts[d] = (int64_t)h[33 + d];
    }
    if (ok) { ctx->n_runs = R; ctx->run_block_bits = bb; ctx->have_runs = true; }
    return BESST_OK;
}

extern "C" int besst_runs_route(besst_ctx* ctx, int32_t world, int64_t* link_counts, int64_t* run_counts) {
    if (!ctx || !ctx->have_runs) { if (ctx) ctx->err = "no grouped runs (besst_links_group)"; return BESST_E_STATE; }
    if (world < 1 || world > 16 || !link_counts || !run_counts) { ctx->err = "runs_route: bad arguments"; return BESST_E_INVALID; }
    cudaSetDevice(ctx->device);
    return besst_launch_runs_route(ctx, world, link_counts, run_counts);
}

extern "C" int besst_runs_pack(besst_ctx* ctx, int32_t world, int32_t* out_obs_device, besst_run_desc* out_desc_device) {
    if (!ctx || !ctx->have_runs) { if (ctx) ctx->err = "no grouped runs (besst_links_group)"; return BESST_E_STATE; }
    if (world < 1 || world > 16 || (ctx->n_runs > 0 && (!out_obs_device || !out_desc_device))) { ctx->err = "runs_pack: bad arguments"; return BESST_E_INVALID; }
    cudaSetDevice(ctx->device);
    return besst_launch_runs_pack(ctx, world, out_obs_device, out_desc_device, nullptr, nullptr);
}

extern "C" int besst_runs_pack_peer(besst_ctx* ctx, int32_t world, int32_t* const* obs_ptrs, besst_run_desc* const* desc_ptrs) {
    if (!ctx || !ctx->have_runs) { if (ctx) ctx->err = "no grouped runs (besst_links_group)"; return BESST_E_STATE; }
    if (world < 1 || world > 16 || !obs_ptrs || !desc_ptrs) { ctx->err = "runs_pack_peer: bad arguments"; return BESST_E_INVALID; }
    cudaSetDevice(ctx->device);
    return besst_launch_runs_pack(ctx, world, nullptr, nullptr, obs_ptrs, desc_ptrs);
}

extern "C" int besst_runs_obs_bytes(const besst_lib_params* params) { return params ? besst_obs_bytes(*params) : -1; }

extern "C" int besst_runs_to_graph(besst_ctx* ctx, const besst_lib_params* params, const int32_t* obs_device, int64_t n_links,
                                   const besst_run_desc* desc_device, int64_t n_runs, int32_t world, int32_t block_bits,
                                   const int64_t* src_run_counts, const int64_t* src_link_counts, const int64_t* src_first_base,
                                   const uint64_t* fishy_keys_device, int64_t n_fishy_keys, besst_graph_sizes* sizes) {
    if (!ctx) return BESST_E_INVALID;
    int rc = check_params(ctx, params);
    if (rc) return rc;
    if (n_links < 0 || n_runs < 0 || n_fishy_keys < 0 || world < 1 || world > 16 || block_bits < 1 || block_bits > 31 ||
        (n_runs > 0 && (!obs_device || !desc_device)) || (n_links > 0 && n_runs == 0) || n_links >= (1ll << 30) ||
        !src_run_counts || !src_link_counts || !src_first_base || (n_fishy_keys > 0 && !fishy_keys_device)) {
        ctx->err = "runs_to_graph: bad arguments";
        return BESST_E_INVALID;
    }
    cudaSetDevice(ctx->device);
    int low_bits = 0;
    rc = besst_launch_runs_import(ctx, desc_device, n_runs, world, block_bits, src_run_counts, src_link_counts, src_first_base, &low_bits);
    if (rc) return rc;
    BesstRunInput in;
    in.grouped = reinterpret_cast<const int2*>(obs_device);
    in.n_runs = n_runs;
    in.low_bits = low_bits;
    in.packed16 = besst_obs_bytes(*params) == 4;
    rc = besst_launch_graph_from_runs(ctx, *params, n_links, in, fishy_keys_device, n_fishy_keys);
    if (rc) return rc;
    ctx->have_runs = false;   // the run buffers now describe the received runs
    if (sizes) {
        sizes->n_edges = ctx->n_edges; sizes->n_links = ctx->n_links; sizes->n_contigs = ctx->n_contigs;
        sizes->n_fishy = n_fishy_keys;
        sizes->n_ll_links = ctx->n_ll_links;
    }
    return BESST_OK;
}

extern "C" int64_t besst_csr_prune_dense(int64_t n_weak, const uint32_t* weak_u, const uint32_t* weak_v, int32_t* degree,
                                         int32_t min_neighbours, uint8_t* dropped) {
    if (n_weak < 0 || (n_weak > 0 && (!weak_u || !weak_v || !degree || !dropped))) return BESST_E_INVALID;
    int64_t removed = 0;
    for (int64_t i = 0; i < n_weak; ++i) {
        const uint32_t u = weak_u[i], v = weak_v[i];
        const bool drop = degree[u] > min_neighbours && degree[v] > min_neighbours;
        dropped[i] = drop ? 1 : 0;
        if (drop) { --degree[u]; --degree[v]; ++removed; }
    }
    return removed;
}

extern "C" int besst_set_stream(besst_ctx* ctx, void* cuda_stream) {
    if (!ctx) return BESST_E_INVALID;
    cudaSetDevice(ctx->device);
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    ctx->use_caller_stream = cuda_stream != nullptr;
    return BESST_OK;
}

// host arrays of n doubles (values) + two length arrays in, one double array out, through ctx->misc
static int batch_f64(besst_ctx* ctx, const besst_lib_params* params, const double* x, const double* len1, const double* len2, int64_t n,
                     double* out, int32_t* gap_out, const char* what,
                     int (*launch)(besst_ctx*, const besst_lib_params&, const double*, const double*, const double*, int64_t, int32_t*, double*)) {
    if (!ctx) return BESST_E_INVALID;
    int rc = check_params(ctx, params);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!x || !len1 || !len2 || (!out && !gap_out)))) { ctx->err = std::string(what) + ": bad arguments"; return BESST_E_INVALID; }
    if (n == 0) return BESST_OK;
    cudaSetDevice(ctx->device);
    const size_t nn = (size_t)n;
    BESST_CUDA_TRY(ctx, ctx->misc.ensure(nn * (4 * 8 + 4) + 64));
    double* d_x = ctx->misc.as<double>();
    double* d_l1 = d_x + nn;
    double* d_l2 = d_l1 + nn;
    double* d_out = d_l2 + nn;
    int32_t* d_gap = reinterpret_cast<int32_t*>(d_out + nn);
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(d_x, x, 8 * nn, cudaMemcpyHostToDevice, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(d_l1, len1, 8 * nn, cudaMemcpyHostToDevice, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(d_l2, len2, 8 * nn, cudaMemcpyHostToDevice, ctx->stream));
    rc = launch(ctx, *params, d_x, d_l1, d_l2, n, d_gap, d_out);
    if (rc) return rc;
    if (gap_out) BESST_CUDA_TRY(ctx, cudaMemcpyAsync(gap_out, d_gap, 4 * nn, cudaMemcpyDeviceToHost, ctx->stream));
    if (out) BESST_CUDA_TRY(ctx, cudaMemcpyAsync(out, d_out, 8 * nn, cudaMemcpyDeviceToHost, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BESST_OK;
}

extern "C" int besst_trsk_sd_batch(besst_ctx* ctx, const besst_lib_params* params, const double* gap, const double* len1,
                                   const double* len2, int64_t n, double* sd_out) {
    if (ctx && n > 0 && !sd_out) { ctx->err = "trsk_sd_batch: bad arguments"; return BESST_E_INVALID; }
    return batch_f64(ctx, params, gap, len1, len2, n, sd_out, nullptr, "trsk_sd_batch",
                     [](besst_ctx* c, const besst_lib_params& p, const double* x, const double* l1, const double* l2, int64_t m, int32_t*, double* o) {
                         return besst_launch_trsk_sd(c, p, x, l1, l2, m, o);
                     });
}

extern "C" int besst_gapest_func_batch(besst_ctx* ctx, const besst_lib_params* params, const double* d, const double* len1,
                                       const double* len2, int64_t n, double* func_out) {
    if (ctx && n > 0 && !func_out) { ctx->err = "gapest_func_batch: bad arguments"; return BESST_E_INVALID; }
    return batch_f64(ctx, params, d, len1, len2, n, func_out, nullptr, "gapest_func_batch",
                     [](besst_ctx* c, const besst_lib_params& p, const double* x, const double* l1, const double* l2, int64_t m, int32_t*, double* o) {
                         return besst_launch_func_of_d(c, p, x, l1, l2, m, o);
                     });
}

extern "C" int besst_links_to_graph(besst_ctx* ctx, const besst_lib_params* params, const besst_link_tuple* tuples_device,
                                    int64_t n_tuples, const uint64_t* fishy_keys_device, int64_t n_fishy_keys,
                                    besst_graph_sizes* sizes) {
    if (!ctx) return BESST_E_INVALID;
    int rc = check_params(ctx, params);
    if (rc) return rc;
    if (n_tuples < 0 || n_fishy_keys < 0 || (n_tuples > 0 && !tuples_device) || (n_fishy_keys > 0 && !fishy_keys_device)) {
        ctx->err = "links_to_graph: bad arguments";
        return BESST_E_INVALID;
    }
    cudaSetDevice(ctx->device);
    rc = besst_launch_graph(ctx, *params, tuples_device, n_tuples, fishy_keys_device, n_fishy_keys);
    if (rc) return rc;
    if (sizes) {
        sizes->n_edges = ctx->n_edges; sizes->n_links = ctx->n_links; sizes->n_contigs = ctx->n_contigs;
        sizes->n_fishy = n_fishy_keys;
        sizes->n_ll_links = ctx->n_ll_links;
    }
    return BESST_OK;
}

extern "C" int besst_graph_build(besst_ctx* ctx, const besst_lib_params* params, const besst_records* records,
                                 besst_graph_sizes* sizes) {
    int64_t n = 0;
    int rc = besst_links_extract(ctx, params, records, &n);
    if (rc) return rc;
    // tuples_device == NULL: the links of this ctx's last extraction, read from their scratch runs
    cudaSetDevice(ctx->device);
    rc = besst_launch_graph(ctx, *params, nullptr, ctx->n_tuples, ctx->fishy_keys.as<uint64_t>(), ctx->n_fishy_keys);
    if (rc) return rc;
    if (sizes) {
        sizes->n_edges = ctx->n_edges; sizes->n_links = ctx->n_links; sizes->n_contigs = ctx->n_contigs;
        sizes->n_fishy = ctx->n_fishy_keys;
        sizes->n_ll_links = ctx->n_ll_links;
    }
    ctx->ev_valid = true;
    return BESST_OK;
}

extern "C" int besst_graph_fetch(besst_ctx* ctx, besst_graph_out* out) {
    if (!ctx || !out) return BESST_E_INVALID;
    if (!ctx->have_graph) { ctx->err = "besst_graph_fetch before a successful build"; return BESST_E_STATE; }
    cudaSetDevice(ctx->device);
    const size_t E = (size_t)ctx->n_edges, L = (size_t)ctx->n_links, C = (size_t)ctx->n_contigs;
#define FETCH(field, buf, bytes)                                                                                      \
    if (out->field && (bytes) > 0)                                                                                   \
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(out->field, ctx->buf.p, (bytes), cudaMemcpyDeviceToHost, ctx->stream))
    FETCH(edge_u, e_u, 4 * E); FETCH(edge_v, e_v, 4 * E); FETCH(nr_links, e_nr, 4 * E);
    FETCH(obs_sum, e_obs, 8 * E); FETCH(obs_sq, e_obs_sq, 8 * E); FETCH(first_idx, e_first, 8 * E);
    FETCH(row_ptr, e_row_ptr, 8 * (E + 1)); FETCH(gap, e_gap, 4 * E); FETCH(score, e_score, 8 * E);
    FETCH(ks, e_ks, 8 * E); FETCH(sd_obs, e_sd_obs, 8 * E); FETCH(sd_model, e_sd_model, 8 * E);
    FETCH(fishy, e_fishy, 4 * E); FETCH(flags, e_flags, E); FETCH(obs_u, l_obs_u, 4 * L); FETCH(obs_v, l_obs_v, 4 * L);
#undef FETCH
    if (out->aligned_len && C > 0 && ctx->have_links)
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(out->aligned_len, ctx->aligned.p, 8 * C, cudaMemcpyDeviceToHost, ctx->stream));
    memset(out->counters, 0, sizeof(out->counters));
    if (ctx->have_links)
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(out->counters, ctx->counters.p, 8 * BESST_N_COUNTERS, cudaMemcpyDeviceToHost, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BESST_OK;
}

extern "C" int besst_graph_view(besst_ctx* ctx, besst_graph_out* out) {
    if (!ctx || !out) return BESST_E_INVALID;
    if (!ctx->have_graph) { ctx->err = "besst_graph_view before a successful build"; return BESST_E_STATE; }
    cudaSetDevice(ctx->device);
    const size_t E = (size_t)ctx->n_edges, L = (size_t)ctx->n_links, C = (size_t)ctx->n_contigs;
    int slot = 0;
#define VIEW(field, type, buf, bytes)                                                                               \
    do {                                                                                                            \
        HBuf& h = ctx->h_out[slot++];                                                                               \
        BESST_CUDA_TRY(ctx, h.ensure((bytes) > 0 ? (bytes) : 1));                                                   \
        if ((bytes) > 0)                                                                                            \
            BESST_CUDA_TRY(ctx, cudaMemcpyAsync(h.p, ctx->buf.p, (bytes), cudaMemcpyDeviceToHost, ctx->stream));    \
        out->field = reinterpret_cast<type*>(h.p);                                                                  \
    } while (0)
    // the two big per-link arrays first: the rest follows while they are on the wire
    VIEW(obs_u, int32_t, l_obs_u, 4 * L); VIEW(obs_v, int32_t, l_obs_v, 4 * L);
    VIEW(edge_u, uint32_t, e_u, 4 * E); VIEW(edge_v, uint32_t, e_v, 4 * E); VIEW(nr_links, int32_t, e_nr, 4 * E);
    VIEW(obs_sum, int64_t, e_obs, 8 * E); VIEW(obs_sq, int64_t, e_obs_sq, 8 * E); VIEW(first_idx, int64_t, e_first, 8 * E);
    VIEW(row_ptr, int64_t, e_row_ptr, 8 * (E + 1)); VIEW(gap, int32_t, e_gap, 4 * E); VIEW(score, double, e_score, 8 * E);
    VIEW(ks, double, e_ks, 8 * E); VIEW(sd_obs, double, e_sd_obs, 8 * E); VIEW(sd_model, double, e_sd_model, 8 * E);
    VIEW(fishy, int32_t, e_fishy, 4 * E); VIEW(flags, uint8_t, e_flags, E);
#undef VIEW
    {
        HBuf& h = ctx->h_out[slot++];
        BESST_CUDA_TRY(ctx, h.ensure(C > 0 ? 8 * C : 1));
        out->aligned_len = reinterpret_cast<int64_t*>(h.p);
        if (C > 0) {
            if (ctx->have_links) BESST_CUDA_TRY(ctx, cudaMemcpyAsync(h.p, ctx->aligned.p, 8 * C, cudaMemcpyDeviceToHost, ctx->stream));
            else memset(h.p, 0, 8 * C);
        }
    }
    memset(out->counters, 0, sizeof(out->counters));
    if (ctx->have_links)
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(out->counters, ctx->counters.p, 8 * BESST_N_COUNTERS, cudaMemcpyDeviceToHost, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BESST_OK;
}

extern "C" int besst_gapest_batch(besst_ctx* ctx, const besst_lib_params* params, const double* mean_obs, const double* len1,
                                  const double* len2, int64_t n, int32_t* gap_out, double* sd_out) {
    if (ctx && n > 0 && !gap_out) { ctx->err = "gapest_batch: bad arguments"; return BESST_E_INVALID; }
    return batch_f64(ctx, params, mean_obs, len1, len2, n, sd_out, gap_out, "gapest_batch",
                     [](besst_ctx* c, const besst_lib_params& p, const double* x, const double* l1, const double* l2, int64_t m, int32_t* g, double* o) {
                         return besst_launch_gapest(c, p, x, l1, l2, m, g, o);
                     });
}

extern "C" int besst_gapest_lognormal_batch(besst_ctx* ctx, double mu_ln, double sigma_ln, double read_len, const int32_t* samples,
                                            const int64_t* row_ptr, const double* len1, const double* len2, int64_t n, int32_t* gap_out) {
    if (!ctx) return BESST_E_INVALID;
    if (n < 0 || (n > 0 && (!samples || !row_ptr || !len1 || !len2 || !gap_out)) || !(sigma_ln > 0)) { ctx->err = "gapest_lognormal_batch: bad arguments"; return BESST_E_INVALID; }
    if (n == 0) return BESST_OK;
    const int64_t total = row_ptr[n];
    for (int64_t i = 0; i < n; ++i)
        if (row_ptr[i + 1] <= row_ptr[i] || row_ptr[i] < 0) { ctx->err = "gapest_lognormal_batch: every edge needs at least one observation"; return BESST_E_INVALID; }
    cudaSetDevice(ctx->device);
    const size_t nn = (size_t)n, tt = (size_t)total;
    BESST_CUDA_TRY(ctx, ctx->misc.ensure(8 * (nn + 1) + 16 * nn + 4 * nn + 4 * tt + 64));
    unsigned char* base = ctx->misc.as<unsigned char>();
    int64_t* d_rp = reinterpret_cast<int64_t*>(base);
    double* d_l1 = reinterpret_cast<double*>(base + 8 * (nn + 1));
    double* d_l2 = d_l1 + nn;
    int32_t* d_gap = reinterpret_cast<int32_t*>(d_l2 + nn);
    int32_t* d_s = d_gap + nn;
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(d_rp, row_ptr, 8 * (nn + 1), cudaMemcpyHostToDevice, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(d_l1, len1, 8 * nn, cudaMemcpyHostToDevice, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(d_l2, len2, 8 * nn, cudaMemcpyHostToDevice, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(d_s, samples, 4 * tt, cudaMemcpyHostToDevice, ctx->stream));
    const int rc = besst_launch_gapest_lognormal(ctx, mu_ln, sigma_ln, read_len, d_s, d_rp, d_l1, d_l2, n, d_gap);
    if (rc) return rc;
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(gap_out, d_gap, 4 * nn, cudaMemcpyDeviceToHost, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BESST_OK;
}

extern "C" int besst_libmetrics(besst_ctx* ctx, const besst_lib_params* params, const besst_records* records,
                                const int64_t* ref_lengths, int64_t n_refs, int32_t want_isize, besst_libmetrics_out* out,
                                double* adjusted_distribution, int64_t cap) {
    if (!ctx) return BESST_E_INVALID;
    int rc = check_params(ctx, params);
    if (rc) return rc;
    if (!out || !ref_lengths || n_refs <= 0) { ctx->err = "libmetrics: bad arguments"; return BESST_E_INVALID; }
    cudaSetDevice(ctx->device);
    DeviceRecords d;
    rc = stage_records(ctx, records, &d, true, false);
    if (rc) return rc;
    return besst_launch_libmetrics(ctx, *params, d, ref_lengths, n_refs, want_isize, out, adjusted_distribution, cap);
}

extern "C" int besst_last_timing(besst_ctx* ctx, float* total_ms, float* stage_ms) {
    if (!ctx || !ctx->ev_valid || ctx->n_stage_marks < 2) { if (ctx) ctx->err = "no timed build"; return BESST_E_STATE; }
    cudaSetDevice(ctx->device);
    BESST_CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev[ctx->n_stage_marks - 1]));
    if (total_ms) BESST_CUDA_TRY(ctx, cudaEventElapsedTime(total_ms, ctx->ev[0], ctx->ev[ctx->n_stage_marks - 1]));
    if (stage_ms)
        for (int i = 0; i < BESST_N_STAGES; ++i) {
            stage_ms[i] = 0.f;
            if (i + 1 < ctx->n_stage_marks) BESST_CUDA_TRY(ctx, cudaEventElapsedTime(&stage_ms[i], ctx->ev[i], ctx->ev[i + 1]));
        }
    return BESST_OK;
}

extern "C" int besst_kernel_launches(besst_ctx* ctx, int64_t* n_launches) {
    if (!ctx || !n_launches) return BESST_E_INVALID;
    *n_launches = ctx->launches;
    return BESST_OK;
}

extern "C" int besst_set_profiling(besst_ctx* ctx, int enabled) {
    if (!ctx) return BESST_E_INVALID;
    ctx->prof = enabled != 0;
    ctx->prof_accumulate = enabled == 2;
    ctx->prof_used = 0;
    return BESST_OK;
}

extern "C" int besst_kernel_profile(besst_ctx* ctx, int32_t* kernel_ids, float* ms, int32_t cap) {
    if (!ctx || !kernel_ids || !ms || cap < 0) return BESST_E_INVALID;
    cudaSetDevice(ctx->device);
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int n = 0;
    for (size_t i = 0; i < ctx->prof_used && n < cap; ++i, ++n) {
        kernel_ids[n] = ctx->prof_pool[i].id;
        BESST_CUDA_TRY(ctx, cudaEventElapsedTime(&ms[n], ctx->prof_pool[i].a, ctx->prof_pool[i].b));
    }
    if (ctx->prof_accumulate) ctx->prof_used = 0;   // read once: the next builds start a new series
    return n;
}
