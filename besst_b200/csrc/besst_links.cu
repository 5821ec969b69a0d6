// K1  records -> accepted link tuples in BAM order.
//
// One pass over the record SoA replaces the reference's per-record Python loop
// (CreateGraph.py:111-211) together with CreateEdge's observation transform,
// duplicate test and acceptance test (CreateGraph.py:812-871, 1024-1076), the
// fishy-pair counting (:141-163, CheckDir :678-688), the coverage accumulation
// (:138-139) and the `counters` bookkeeping (Parameter.py:113-124).
//
// Design (HBM-bound streaming kernel, no tensor cores):
//  * persistent CTAs (grid = SMs x resident CTAs), tiles of 1024 records handed
//    out by an atomic ticket so that tile t-1 is always running when t starts;
//  * 128-bit coalesced loads of the SoA columns (4 consecutive records/thread);
//    the 32-byte contig rows are two 128-bit gathers that hit L1/L2 because the
//    BAM is tid-sorted;
//  * the only order-dependent state of the reference -- "previous CreateEdge
//    call's (obs1,obs2)" -- is a rightmost-non-empty scan, done with warp
//    shuffles inside the tile and a decoupled look-back across tiles
//    (self-validating 64-bit words, no fences);
//  * accepted tuples are compacted in BAM order (second decoupled look-back on
//    the counts), staged in shared memory and written with 16-byte stores;
//  * coverage uses __match_any_sync warp-aggregated 64-bit atomics; counters
//    stay in registers for the life of the CTA.
#include "besst_internal.cuh"

namespace {

constexpr int K1_THREADS = 256;
constexpr int K1_ITEMS = 4;
constexpr int K1_TILE = K1_THREADS * K1_ITEMS;
constexpr int K1_WARPS = K1_THREADS / 32;

typedef unsigned long long u64;

struct Row {
    int4 a;  // state, scaffold, direction, position
    int4 b;  // length, scaf_length, in_largest, reserved
};

struct K1Params {
    DeviceRecords rec;
    const Row* rows;
    int n_contigs;
    int orientation, min_mapq, detect_dup, extend, scoring;
    double read_len, threshold;
    int halo1, halo2;
    besst_link_tuple* out;
    long long out_cap;
    u64* fishy;
    long long fishy_cap;
    u64* aligned;
    u64* counters;  // [BESST_N_COUNTERS]
    u64* globals;   // [0]=ticket [1]=n_out [2]=n_fishy
    u64 *eligA0, *eligA1, *eligP0, *eligP1, *acc;
    int n_tiles;
};

constexpr u64 READY = 1ull << 63;
constexpr u64 HAS = 1ull << 62;
constexpr u64 ST_AGG = 1ull << 62;
constexpr u64 ST_INC = 2ull << 62;
constexpr u64 ST_MASK = 3ull << 62;

struct Last {
    int has, o1, o2;
};

__device__ __forceinline__ u64 ld_vol(const u64* p) { return *reinterpret_cast<const volatile u64*>(p); }
__device__ __forceinline__ void st_vol(u64* p, u64 v) { *reinterpret_cast<volatile u64*>(p) = v; }

__device__ __forceinline__ Last warp_scan_last(Last v, int lane) {
    // inclusive scan of "rightmost non-empty"
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int h = __shfl_up_sync(0xffffffffu, v.has, off);
        int a = __shfl_up_sync(0xffffffffu, v.o1, off);
        int b = __shfl_up_sync(0xffffffffu, v.o2, off);
        if (lane >= off && !v.has) {
            v.has = h;
            v.o1 = a;
            v.o2 = b;
        }
    }
    return v;
}

// PosDirCalculatorPE / PosDirCalculatorMP (CreateGraph.py:1024-1076) for one end
__device__ __forceinline__ void pos_dir(int cdir, int read_fwd, int orientation, long long cpos, long long rpos,
                                        long long slen, long long clen, double read_len, int& obs, int& side_r) {
    int fwd = orientation == BESST_ORIENT_FR ? read_fwd : !read_fwd;
    double o;
    if (cdir && fwd) {
        o = (double)(slen - cpos - rpos);
        side_r = 1;
    } else if (!cdir && fwd) {
        o = (double)(cpos + (clen - rpos));
        side_r = 0;
    } else if (cdir && !fwd) {
        o = __dadd_rn((double)(cpos + rpos), read_len);
        side_r = 0;
    } else {
        o = __dsub_rn((double)(slen - cpos), __dsub_rn((double)(clen - rpos), read_len));
        side_r = 1;
    }
    obs = __double2int_rz(o);
}

template <bool VEC>
__device__ __forceinline__ void load_i32x4(const int32_t* p, long long idx, long long n, int (&v)[K1_ITEMS]) {
    if (VEC) {
        int4 t = __ldg(reinterpret_cast<const int4*>(p + idx));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int i = 0; i < K1_ITEMS; ++i) v[i] = (idx + i < n) ? __ldg(p + idx + i) : -1;
    }
}

template <bool VEC>
__global__ void __launch_bounds__(K1_THREADS) k_extract_links(const K1Params P) {
    __shared__ int s_tile;
    __shared__ Last s_warp_last[K1_WARPS];
    __shared__ Last s_pred;
    __shared__ int s_warp_cnt[K1_WARPS];
    __shared__ long long s_base;
    __shared__ u64 s_cnt[8];
    __shared__ __align__(16) besst_link_tuple s_out[K1_TILE];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int c_count = 0, c_nonuniq = 0, c_nonuniq_scaf = 0, c_dups = 0, c_toolong = 0, c_fishy = 0, c_calls = 0, c_valid = 0;
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;

    for (;;) {
        __syncthreads();  // protects s_tile / s_out / scan scratch of the previous tile
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(&P.globals[0], 1ull);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= P.n_tiles) break;
        const long long idx0 = (long long)tile * K1_TILE + (long long)threadIdx.x * K1_ITEMS;
        const long long n = P.rec.n;
        const bool full = VEC && ((long long)(tile + 1) * K1_TILE <= n);

        int tid[K1_ITEMS], mtid[K1_ITEMS], qlen[K1_ITEMS];
        unsigned flag[K1_ITEMS], mapq[K1_ITEMS];
        if (full) {
            load_i32x4<true>(P.rec.tid, idx0, n, tid);
            load_i32x4<true>(P.rec.mtid, idx0, n, mtid);
            load_i32x4<true>(P.rec.qlen, idx0, n, qlen);
            uint2 f = __ldg(reinterpret_cast<const uint2*>(P.rec.flag + idx0));
            flag[0] = f.x & 0xffffu; flag[1] = f.x >> 16; flag[2] = f.y & 0xffffu; flag[3] = f.y >> 16;
            unsigned m = __ldg(reinterpret_cast<const unsigned*>(P.rec.mapq + idx0));
            mapq[0] = m & 0xffu; mapq[1] = (m >> 8) & 0xffu; mapq[2] = (m >> 16) & 0xffu; mapq[3] = m >> 24;
        } else {
            load_i32x4<false>(P.rec.tid, idx0, n, tid);
            load_i32x4<false>(P.rec.mtid, idx0, n, mtid);
            load_i32x4<false>(P.rec.qlen, idx0, n, qlen);
#pragma unroll
            for (int i = 0; i < K1_ITEMS; ++i) {
                flag[i] = (idx0 + i < n) ? __ldg(P.rec.flag + idx0 + i) : 0u;
                mapq[i] = (idx0 + i < n) ? __ldg(P.rec.mapq + idx0 + i) : 0u;
            }
        }

        // ---- per-record classification ------------------------------------
        bool elig[K1_ITEMS], cov[K1_ITEMS], ll[K1_ITEMS];
        int o1[K1_ITEMS], o2[K1_ITEMS];
        unsigned nu[K1_ITEMS], nv[K1_ITEMS];
        bool need_pos = false;
        int4 r1a[K1_ITEMS], r2a[K1_ITEMS];  // state, scaffold, direction, position
#pragma unroll
        for (int i = 0; i < K1_ITEMS; ++i) {
            elig[i] = false; cov[i] = false; ll[i] = false;
            o1[i] = o2[i] = 0; nu[i] = nv[i] = 0;
            bool ok = tid[i] >= 0 && mtid[i] >= 0 && tid[i] < P.n_contigs && mtid[i] < P.n_contigs;  // :118-124
            if (ok) {
                r1a[i] = __ldg(reinterpret_cast<const int4*>(P.rows + tid[i]));
                r2a[i] = __ldg(reinterpret_cast<const int4*>(P.rows + mtid[i]));
                ok = r1a[i].x != BESST_CTG_ABSENT && r2a[i].x != BESST_CTG_ABSENT;                       // :127-130
            }
            if (!ok) continue;
            c_valid++;
            const unsigned f = flag[i];
            const bool unmapped = f & 0x4u, read1 = f & 0x40u, read2 = f & 0x80u;
            const int mq = (int)mapq[i];
            cov[i] = (mq >= P.min_mapq) || mq == 0;                                                      // :138
            const bool diff_scaf = r1a[i].y != r2a[i].y;
            if (unmapped && read1 && diff_scaf) {                                                        // :141-163
                int d1, d2, s1, s2;
                pos_dir(r1a[i].z, !(f & 0x10u), P.orientation, 0, 0, 0, 0, 0.0, d1, s1);
                pos_dir(r2a[i].z, !(f & 0x20u), P.orientation, 0, 0, 0, 0, 0.0, d2, s2);
                unsigned n1 = 2u * (unsigned)r1a[i].y + (unsigned)s1, n2 = 2u * (unsigned)r2a[i].y + (unsigned)s2;
                u64 key = n1 < n2 ? (((u64)n1 << 32) | n2) : (((u64)n2 << 32) | n1);
                u64 slot = atomicAdd(&P.globals[2], 1ull);
                if ((long long)slot < P.fishy_cap) P.fishy[slot] = key;
                c_fishy++;
            }
            const bool inter = tid[i] != mtid[i];
            if (inter && mq == 0) c_nonuniq++;                                                           // :166-167
            if (inter && read2 && !unmapped && mq >= P.min_mapq) {                                       // :169
                const bool l1 = r1a[i].x == BESST_CTG_LARGE, l2 = r2a[i].x == BESST_CTG_LARGE;
                if (l1 && l2) {
                    if (diff_scaf) { elig[i] = true; ll[i] = true; }                                     // :170
                } else if (P.extend) {                                                                   // :184-206
                    if (!(l1 || l2) ? diff_scaf : true) elig[i] = true;
                }
            }
            need_pos |= elig[i];
        }

        if (__any_sync(0xffffffffu, need_pos)) {
            int pos[K1_ITEMS], mpos[K1_ITEMS];
            if (full) {
                load_i32x4<true>(P.rec.pos, idx0, n, pos);
                load_i32x4<true>(P.rec.mpos, idx0, n, mpos);
            } else {
                load_i32x4<false>(P.rec.pos, idx0, n, pos);
                load_i32x4<false>(P.rec.mpos, idx0, n, mpos);
            }
#pragma unroll
            for (int i = 0; i < K1_ITEMS; ++i) {
                if (!elig[i]) continue;
                int s1, s2;
                const unsigned f = flag[i];
                // length, scaf_length live in the second half of the row: only link records need them
                const int4 r1b = __ldg(reinterpret_cast<const int4*>(P.rows + tid[i]) + 1);
                const int4 r2b = __ldg(reinterpret_cast<const int4*>(P.rows + mtid[i]) + 1);
                pos_dir(r1a[i].z, !(f & 0x10u), P.orientation, r1a[i].w, pos[i], r1b.y, r1b.x, P.read_len, o1[i], s1);
                pos_dir(r2a[i].z, !(f & 0x20u), P.orientation, r2a[i].w, mpos[i], r2b.y, r2b.x, P.read_len, o2[i], s2);
                nu[i] = 2u * (unsigned)r1a[i].y + (unsigned)s1;
                nv[i] = 2u * (unsigned)r2a[i].y + (unsigned)s2;
            }
        }

        // ---- coverage: warp-aggregated 64-bit atomics (:138-139) ------------
        {
            int t0 = -1, s0 = 0;
#pragma unroll
            for (int i = 0; i < K1_ITEMS; ++i)
                if (cov[i]) {
                    if (t0 < 0) t0 = tid[i];
                    if (tid[i] == t0) { s0 += qlen[i]; cov[i] = false; }
                }
            unsigned act = __ballot_sync(0xffffffffu, t0 >= 0);
            if (t0 >= 0) {
                unsigned peers = __match_any_sync(act, t0);
                int sum = __reduce_add_sync(peers, s0);
                if (lane == __ffs(peers) - 1) atomicAdd(&P.aligned[t0], (u64)(long long)sum);
            }
#pragma unroll
            for (int i = 1; i < K1_ITEMS; ++i) {   // records of a second/third contig inside one thread: rare
                unsigned rest = __ballot_sync(0xffffffffu, cov[i]);
                if (rest && cov[i]) {
                    unsigned peers = __match_any_sync(rest, tid[i]);
                    int sum = __reduce_add_sync(peers, qlen[i]);
                    if (lane == __ffs(peers) - 1) atomicAdd(&P.aligned[tid[i]], (u64)(long long)sum);
                }
            }
        }

        // ---- previous CreateEdge call: rightmost-non-empty scan -------------
        Last mine = {0, 0, 0};
#pragma unroll
        for (int i = 0; i < K1_ITEMS; ++i)
            if (elig[i]) { mine.has = 1; mine.o1 = o1[i]; mine.o2 = o2[i]; }
        Last incl = warp_scan_last(mine, lane);
        Last excl;
        excl.has = __shfl_up_sync(0xffffffffu, incl.has, 1);
        excl.o1 = __shfl_up_sync(0xffffffffu, incl.o1, 1);
        excl.o2 = __shfl_up_sync(0xffffffffu, incl.o2, 1);
        if (lane == 0) excl.has = 0;
        if (lane == 31) s_warp_last[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            Last tot = {0, 0, 0};
            for (int w = 0; w < K1_WARPS; ++w)
                if (s_warp_last[w].has) tot = s_warp_last[w];
            // publish the tile-local value, then resolve the predecessor
            st_vol(&P.eligA1[tile], READY | (u64)(unsigned)tot.o2);
            st_vol(&P.eligA0[tile], READY | (tot.has ? HAS : 0) | (u64)(unsigned)tot.o1);
            Last pred = {0, 0, 0};
            int p = tile - 1;
            while (p >= 0) {
                u64 w0 = ld_vol(&P.eligP0[p]);
                if (w0 & READY) {  // inclusive value of everything up to p
                    u64 w1;
                    do { w1 = ld_vol(&P.eligP1[p]); } while (!(w1 & READY));
                    pred.has = (w0 & HAS) ? 1 : 0; pred.o1 = (int)(unsigned)w0; pred.o2 = (int)(unsigned)w1;
                    break;
                }
                w0 = ld_vol(&P.eligA0[p]);
                if (!(w0 & READY)) continue;  // spin on tile p
                if (w0 & HAS) {
                    u64 w1;
                    do { w1 = ld_vol(&P.eligA1[p]); } while (!(w1 & READY));
                    pred.has = 1; pred.o1 = (int)(unsigned)w0; pred.o2 = (int)(unsigned)w1;
                    break;
                }
                --p;
            }
            Last inc = tot.has ? tot : pred;
            st_vol(&P.eligP1[tile], READY | (u64)(unsigned)inc.o2);
            st_vol(&P.eligP0[tile], READY | (inc.has ? HAS : 0) | (u64)(unsigned)inc.o1);
            if (tile == P.n_tiles - 1) {  // halo for the next rank
                P.counters[BESST_CNT_LAST_OBS1] = (u64)(long long)(inc.has ? inc.o1 : P.halo1);
                P.counters[BESST_CNT_LAST_OBS2] = (u64)(long long)(inc.has ? inc.o2 : P.halo2);
            }
            s_pred = pred;
        }
        __syncthreads();
        Last prev = s_pred;
        for (int w = 0; w < warp; ++w)
            if (s_warp_last[w].has) prev = s_warp_last[w];
        if (excl.has) prev = excl;
        int p1 = prev.has ? prev.o1 : P.halo1, p2 = prev.has ? prev.o2 : P.halo2;

        // ---- CreateEdge: duplicate test, acceptance test, counters (:835-870)
        bool acc[K1_ITEMS];
        int n_acc = 0;
#pragma unroll
        for (int i = 0; i < K1_ITEMS; ++i) {
            acc[i] = false;
            if (!elig[i]) continue;
            c_calls++;
            const bool mq0 = mapq[i] == 0;
            if (mq0) c_nonuniq_scaf++;
            const bool dup = (o1[i] == p1 && o2[i] == p2);
            p1 = o1[i]; p2 = o2[i];
            bool is_dupl = false;
            if (dup) { c_dups++; is_dupl = P.detect_dup; }
            const bool pass = (double)((long long)o1[i] + o2[i]) < P.threshold && o1[i] > 25 && o2[i] > 25;
            if (!is_dupl) {
                if (pass) { c_count++; acc[i] = true; n_acc++; } else c_toolong++;
            }
            if (ll[i] && P.extend && P.scoring && !is_dupl) {  // second call into G_prime (:180-183)
                if (mq0) c_nonuniq_scaf++;
                const bool dup2 = (o1[i] == -1 && o2[i] == -1);
                if (dup2) c_dups++;
                if (!(dup2 && P.detect_dup)) { if (pass) c_count++; else c_toolong++; }
            }
        }

        // ---- ordered compaction of the accepted tuples ----------------------
        int incl_cnt = n_acc;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl_cnt, off);
            if (lane >= off) incl_cnt += t;
        }
        if (lane == 31) s_warp_cnt[warp] = incl_cnt;
        __syncthreads();
        int tile_acc = 0, warp_off = 0;
#pragma unroll
        for (int w = 0; w < K1_WARPS; ++w) {
            if (w < warp) warp_off += s_warp_cnt[w];
            tile_acc += s_warp_cnt[w];
        }
        if (warp == 0) {  // decoupled look-back on the accepted counts
            if (lane == 0) st_vol(&P.acc[tile], (tile == 0 ? ST_INC : ST_AGG) | (u64)tile_acc);
            long long excl_sum = 0;
            if (tile > 0) {
                int p = tile - 1 - lane;
                for (;;) {
                    u64 w;
                    unsigned ready;
                    do {
                        w = (p >= 0) ? ld_vol(&P.acc[p]) : ST_INC;
                        ready = __ballot_sync(0xffffffffu, (w & ST_MASK) != 0);
                    } while (ready != 0xffffffffu);
                    unsigned inc_mask = __ballot_sync(0xffffffffu, (w & ST_MASK) == ST_INC);
                    long long v = (long long)(w & ~ST_MASK);
                    if (inc_mask) {
                        int first = __ffs(inc_mask) - 1;
                        if (lane > first) v = 0;
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    excl_sum += v;
                    if (inc_mask) break;
                    p -= 32;
                }
                if (lane == 0) st_vol(&P.acc[tile], ST_INC | (u64)(excl_sum + tile_acc));
            }
            if (lane == 0) {
                s_base = excl_sum;
                if (tile == P.n_tiles - 1) P.globals[1] = (u64)(excl_sum + tile_acc);
            }
        }
        int local = warp_off + incl_cnt - n_acc;
#pragma unroll
        for (int i = 0; i < K1_ITEMS; ++i)
            if (acc[i]) {
                besst_link_tuple t;
                if (nu[i] < nv[i]) { t.u = nu[i]; t.v = nv[i]; t.obs_u = o1[i]; t.obs_v = o2[i]; }
                else { t.u = nv[i]; t.v = nu[i]; t.obs_u = o2[i]; t.obs_v = o1[i]; }
                s_out[local++] = t;
            }
        __syncthreads();
        const long long base = s_base;
        const int4* src = reinterpret_cast<const int4*>(s_out);
        int4* dst = reinterpret_cast<int4*>(P.out);
        for (int j = threadIdx.x; j < tile_acc; j += K1_THREADS)
            if (base + j < P.out_cap) dst[base + j] = src[j];
    }

    // ---- flush the per-CTA counters -------------------------------------------
    __syncthreads();
    const int local_cnt[8] = {c_count, c_nonuniq, c_nonuniq_scaf, c_dups, c_toolong, c_fishy, c_calls, c_valid};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int v = local_cnt[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0 && v) atomicAdd(&s_cnt[k], (u64)(long long)v);
    }
    __syncthreads();
    if (threadIdx.x < 8 && s_cnt[threadIdx.x]) atomicAdd(&P.counters[threadIdx.x], s_cnt[threadIdx.x]);
}

}  // namespace

int besst_launch_extract(besst_ctx* ctx, const besst_lib_params& p, const DeviceRecords& rec) {
    const int64_t n = rec.n;
    const int64_t n_tiles64 = (n + K1_TILE - 1) / K1_TILE;
    if (n_tiles64 > 0x7fffffff) { ctx->err = "too many records for one call"; return BESST_E_INVALID; }
    const int n_tiles = (int)n_tiles64;
    BESST_CUDA_TRY(ctx, ctx->aligned.ensure(sizeof(u64) * (size_t)(ctx->n_contigs + 1)));
    BESST_CUDA_TRY(ctx, ctx->counters.ensure(sizeof(u64) * (BESST_N_COUNTERS + 8)));
    BESST_CUDA_TRY(ctx, ctx->tile_state.ensure(sizeof(u64) * 5 * (size_t)(n_tiles + 1)));
    if (ctx->tuples_cap == 0) ctx->tuples_cap = n / 2 + 4096;
    if (ctx->fishy_cap == 0) ctx->fishy_cap = n / 8 + 4096;

    bool vec = true;
    const void* ptrs[] = {rec.tid, rec.mtid, rec.pos, rec.mpos, rec.qlen, rec.flag, rec.mapq};
    for (const void* q : ptrs) vec = vec && ((reinterpret_cast<uintptr_t>(q) & 15u) == 0);

    int per_sm = 0;
    if (vec) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_extract_links<true>, K1_THREADS, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_extract_links<false>, K1_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    int grid = ctx->sm_count * per_sm;
    if (grid > n_tiles) grid = n_tiles > 0 ? n_tiles : 1;

    for (int attempt = 0; attempt < 3; ++attempt) {
        BESST_CUDA_TRY(ctx, ctx->tuples.ensure(sizeof(besst_link_tuple) * (size_t)ctx->tuples_cap));
        BESST_CUDA_TRY(ctx, ctx->fishy_keys.ensure(sizeof(u64) * (size_t)ctx->fishy_cap));
        BESST_CUDA_TRY(ctx, cudaMemsetAsync(ctx->aligned.p, 0, sizeof(u64) * (size_t)(ctx->n_contigs + 1), ctx->stream));
        BESST_CUDA_TRY(ctx, cudaMemsetAsync(ctx->counters.p, 0, sizeof(u64) * (BESST_N_COUNTERS + 8), ctx->stream));
        BESST_CUDA_TRY(ctx, cudaMemsetAsync(ctx->tile_state.p, 0, sizeof(u64) * 5 * (size_t)(n_tiles + 1), ctx->stream));
        K1Params P;
        P.rec = rec;
        P.rows = ctx->rows.as<Row>();
        P.n_contigs = (int)ctx->n_contigs;
        P.orientation = p.orientation; P.min_mapq = p.min_mapq; P.detect_dup = p.detect_duplicate;
        P.extend = p.extend_paths; P.scoring = !p.no_score;
        P.read_len = p.read_len; P.threshold = p.ins_size_threshold;
        P.halo1 = p.halo_prev_obs1; P.halo2 = p.halo_prev_obs2;
        P.out = ctx->tuples.as<besst_link_tuple>(); P.out_cap = ctx->tuples_cap;
        P.fishy = ctx->fishy_keys.as<u64>(); P.fishy_cap = ctx->fishy_cap;
        P.aligned = ctx->aligned.as<u64>();
        P.counters = ctx->counters.as<u64>();
        P.globals = ctx->counters.as<u64>() + BESST_N_COUNTERS;
        u64* ts = ctx->tile_state.as<u64>();
        const size_t stride = (size_t)(n_tiles + 1);
        P.eligA0 = ts; P.eligA1 = ts + stride; P.eligP0 = ts + 2 * stride; P.eligP1 = ts + 3 * stride; P.acc = ts + 4 * stride;
        P.n_tiles = n_tiles;
        if (n_tiles > 0) {
            KTimer kt(ctx, BESST_K_EXTRACT);
            if (vec) k_extract_links<true><<<grid, K1_THREADS, 0, ctx->stream>>>(P);
            else k_extract_links<false><<<grid, K1_THREADS, 0, ctx->stream>>>(P);
            BESST_CUDA_TRY(ctx, cudaGetLastError());
        }
        u64 g[3] = {0, 0, 0};
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(g, P.globals, sizeof(g), cudaMemcpyDeviceToHost, ctx->stream));
        BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (n_tiles == 0) {
            int64_t halo[2] = {p.halo_prev_obs1, p.halo_prev_obs2};
            BESST_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->counters.as<u64>() + BESST_CNT_LAST_OBS1, halo, sizeof(halo),
                                                cudaMemcpyHostToDevice, ctx->stream));
        }
        ctx->n_tuples = (int64_t)g[1];
        ctx->n_fishy_keys = (int64_t)g[2];
        bool again = false;
        if (ctx->n_tuples > ctx->tuples_cap) { ctx->tuples_cap = ctx->n_tuples + 4096; again = true; }
        if (ctx->n_fishy_keys > ctx->fishy_cap) { ctx->fishy_cap = ctx->n_fishy_keys + 4096; again = true; }
        if (!again) { ctx->have_links = true; return BESST_OK; }
    }
    ctx->err = "link extraction: capacity retry failed";
    return BESST_E_STATE;
}
