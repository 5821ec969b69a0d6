// K1  records -> accepted link tuples in BAM order.
//
// Replaces the reference's per-record Python loop (CreateGraph.py:111-211)
// together with CreateEdge's observation transform, duplicate test and
// acceptance test (CreateGraph.py:812-871, 1024-1076), the fishy-pair counting
// (:141-163, CheckDir :678-688), the coverage accumulation (:138-139) and the
// `counters` bookkeeping (Parameter.py:113-124).
//
// Design (HBM-bound streaming, no tensor cores, no inter-CTA waiting):
//
//  k_extract_links_tma   one WARP owns a tile of 128 consecutive records, staged in shared
//      memory by TMA bulk copies (cp.async.bulk + mbarrier), double buffered per warp; the contig
//      table is a packed 16-byte row, one 128-bit gather per read end, L1/L2 resident because the
//      BAM is tid-sorted.  No block barrier in the main loop.  The only order-dependent
//      state of the reference -- "(obs1,obs2) of the previous CreateEdge call"
//      -- is a rightmost-non-empty scan: inside the tile it is done with warp
//      shuffles; across tiles only the FIRST eligible record of a tile depends
//      on earlier tiles, so the tile assumes "not a duplicate", writes its
//      accepted tuples compacted into a tile-local slot of a scratch array and
//      publishes a 48-byte aggregate (first/last eligible observation, count,
//      what to undo if the first one turns out to be a duplicate).
//  k_tile_reduce / k_chunk_resolve / k_tile_offsets   a two-level scan over the
//      aggregates (1024 tiles per chunk) resolves every tile boundary: duplicate
//      verdict, counter corrections, exact output offset.  O(N/128) work.
//  k_compact_tuples  copies each tile's run to its final position (16-byte
//      loads/stores), dropping a boundary duplicate.
//
// Coverage uses warp-aggregated 64-bit atomics; counters stay
// in registers for the life of a warp and are flushed once.
#include <math.h>
#include <stdlib.h>

#include "besst_internal.cuh"

namespace {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int WT_ITEMS = 4;
constexpr int WT = 32 * WT_ITEMS;   // records per warp tile
constexpr int SC_THREADS = 1024;    // aggregates per scan chunk

constexpr u32 AGG_HAS = 1u << 31;     // the tile holds at least one CreateEdge call
constexpr u32 AGG_PASS = 1u << 30;    // its first call passes the acceptance test (:840)
constexpr u32 AGG_SECOND = 1u << 29;  // ... and is followed by the second call into G_prime (:180-183)
constexpr u32 AGG_MQ0 = 1u << 28;     // ... on a mapq-0 record (non_unique_for_scaf, :814-815)

// Aggregate of a run of records (a tile, or a chunk of tiles): same shape at both scan levels.
struct __align__(16) Agg {
    u32 flags;
    int first_o1, first_o2;   // (obs1,obs2) of the first CreateEdge call in the run
    int last_o1, last_o2;     // ... of the last one
    u32 cnt;                  // accepted tuples, the run's first call assumed not to be a duplicate
    int dups, count_corr, toolong_corr, nus_corr;   // counter corrections already resolved inside the run
    u32 pad0, pad1;
};
static_assert(sizeof(Agg) == 48, "Agg layout");

struct ChunkIn {   // what a chunk receives from everything before it
    int carry_o1, carry_o2;
    u64 offset;
};

constexpr u64 OFF_DROP = 1ull << 63;   // tile_off flag: skip the tile's first scratch tuple
constexpr u64 OFF_MASK = OFF_DROP - 1;

struct K1Params {
    DeviceRecords rec;
    const int4* rows;   // packed contig rows
    int n_contigs;
    int orientation, min_mapq, detect_dup, extend, scoring;
    double read_len;
    int read_len_i;          // read_len when it is integral
    long long threshold_i;   // ceil(ins_size_threshold): obs1 + obs2 < threshold  <=>  obs1 + obs2 < threshold_i
    besst_link_tuple* scratch;   // [n_tiles * WT]
    Agg* aggs;                   // [n_tiles]
    u64* fishy;
    long long fishy_cap;
    u64* aligned;
    u64* counters;  // [BESST_N_COUNTERS]
    u64* globals;   // [0]=n_out [1]=n_fishy
    long long n_tiles;
};

struct Last {
    int has, o1, o2;
};

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

// PosDirCalculatorPE / PosDirCalculatorMP (CreateGraph.py:1024-1076) for one end.  The reference
// adds a possibly fractional read_len in fp64 and truncates with int(); when read_len is integral
// (INT_RL) every intermediate is an exact integer and the fp64 round trip is skipped.
template <bool INT_RL>
__device__ __forceinline__ void pos_dir(int cdir, int read_fwd, int orientation, int cpos, int rpos, int slen, int clen,
                                        double read_len, int read_len_i, int& obs, int& side_r) {
    const int fwd = orientation == BESST_ORIENT_FR ? read_fwd : !read_fwd;
    if (cdir && fwd) {
        obs = slen - cpos - rpos;
        side_r = 1;
    } else if (!cdir && fwd) {
        obs = cpos + (clen - rpos);
        side_r = 0;
    } else if (cdir && !fwd) {
        if (INT_RL) obs = cpos + rpos + read_len_i;
        else obs = __double2int_rz(__dadd_rn((double)(cpos + rpos), read_len));
        side_r = 0;
    } else {
        if (INT_RL) obs = (slen - cpos) - ((clen - rpos) - read_len_i);
        else obs = __double2int_rz(__dsub_rn((double)(slen - cpos), __dsub_rn((double)(clen - rpos), read_len)));
        side_r = 1;
    }
}

// sides only (CheckDir, CreateGraph.py:678-688)
__device__ __forceinline__ int side_only(int cdir, int read_fwd, int orientation) {
    int fwd = orientation == BESST_ORIENT_FR ? read_fwd : !read_fwd;
    return (cdir && fwd) || (!cdir && !fwd);
}

__device__ __forceinline__ int row_state(int x) { return x & 3; }
__device__ __forceinline__ int row_dir(int x) { return (x >> 2) & 1; }
__device__ __forceinline__ int row_scaf(int x) { return (int)((u32)x >> 3); }

// ---- K1 -------------------------------------------------------------------------------------------
// How the record columns reach the SM and how the per-record work is laid out:
//  * every warp owns two shared-memory stages of one 128-record tile (7 columns, 2944 B).  Lane 0
//    issues the tile after next with cp.async.bulk (TMA, 1-D) completing on a per-stage mbarrier,
//    so a warp always has a whole tile in flight while it works on the current one and no register
//    is tied up by outstanding loads.  pos/mpos are only fetched ahead when the previous tile had a
//    CreateEdge candidate (libraries with long contigs keep reading 15 instead of 23 B/record); a
//    tile that needs them unexpectedly fetches them on a third mbarrier.
//  * candidates are compacted as one-byte record indices into the staged tile instead of copying
//    five words each.
//  * the contig-table word of a lane's first record is reused for its other three records (the BAM
//    is tid-sorted); out-of-range contigs read as row word 0 == BESST_CTG_ABSENT.
// tuning knobs (overridable with -D for A/B builds, scripts/variants.sh)
#ifndef BESST_K1T_WARPS
#define BESST_K1T_WARPS 8
#endif
#ifndef BESST_K1T_MIN_CTAS
#define BESST_K1T_MIN_CTAS 4
#endif
#ifndef BESST_K1T_BATCH
#define BESST_K1T_BATCH 8
#endif
constexpr int K1T_WARPS = BESST_K1T_WARPS;
constexpr int K1T_THREADS = 32 * K1T_WARPS;
constexpr int K1T_MIN_CTAS = BESST_K1T_MIN_CTAS;
constexpr int K1T_BATCH = BESST_K1T_BATCH;   // consecutive tiles per ticket

struct __align__(16) TileBuf {
    int tid[WT], mtid[WT], qlen[WT], pos[WT], mpos[WT];
    unsigned short flag[WT];
    unsigned char mapq[WT];
};
static_assert(sizeof(TileBuf) == 2944, "TileBuf layout");
constexpr u32 TILE_BYTES_NOPOS = 3 * 4 * WT + 2 * WT + WT;   // tid, mtid, qlen, flag, mapq
constexpr u32 TILE_BYTES_POS = 2 * 4 * WT;

struct __align__(16) WarpSmem {
    TileBuf buf[2];
    u64 bar[4];                 // [0],[1]: stage barriers, [2]: on-demand pos/mpos
    unsigned char cand[WT];     // record index (within the tile) of every CreateEdge candidate, BAM order
};
static_assert(sizeof(WarpSmem) % 16 == 0, "WarpSmem alignment");

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    u32 pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    u32 done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

// PACKED: flag | mapq << 12 | qlen << 20 arrive as ONE 32-bit column (besst_records.packed) staged where the qlen
// column would be: 20 instead of 23 bytes per record over PCIe and out of HBM
template <bool INT_RL, bool PACKED>
__global__ void __launch_bounds__(K1T_THREADS, K1T_MIN_CTAS) k_extract_links_tma(const K1Params P, const int vec_ok) {
    extern __shared__ __align__(128) unsigned char k1_smem[];
    __shared__ u64 s_cnt[8];
    const int lane = threadIdx.x & 31;
    // broadcast from lane 0 so that the compiler knows the warp index (and everything derived from it:
    // tile numbers, stage addresses) is warp-uniform and keeps the TMA operands in uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const u32 lt_mask = (1u << lane) - 1u;
    WarpSmem& W = reinterpret_cast<WarpSmem*>(k1_smem)[warp];
    const u32 bar0 = smem_u32(&W.bar[0]);
    int c_count = 0, c_nonuniq = 0, c_nonuniq_scaf = 0, c_dups = 0, c_toolong = 0, c_fishy = 0, c_calls = 0, c_valid = 0;
    int c_postiles = 0;   // warp-uniform: tiles whose pos / mpos columns were fetched (the kernel's real input bytes)
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    if (lane == 0) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_init(bar0 + 16, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long n = P.rec.n;
    const int warp_global = (int)blockIdx.x * K1T_WARPS + warp;
    const int n_warps = (int)gridDim.x * K1T_WARPS;
    const u32 n_contigs = (u32)P.n_contigs;
    // acceptance test obs1 + obs2 < ins_size_threshold (:840) in 32 bits: both terms are > 25 when it matters
    const bool thr_all = P.threshold_i > 0xffffffffll;
    const u32 thr_u = P.threshold_i <= 0 ? 0u : (thr_all ? 0xffffffffu : (u32)P.threshold_i);
    const bool second_on = P.extend && P.scoring;
    u32 demand_par = 0;

    // ---- one tile, staged in B.  pos_ready: its pos/mpos columns are (being) loaded.  -> it had a candidate
    auto process = [&](const long long wt, TileBuf& B, const bool pos_ready, const bool can_demand) -> bool {
        int tid[WT_ITEMS], mtid[WT_ITEMS], qlen[WT_ITEMS];
        u32 flag[WT_ITEMS], mapq[WT_ITEMS];
        {
            const int4 a = *reinterpret_cast<const int4*>(&B.tid[lane * WT_ITEMS]);
            const int4 b = *reinterpret_cast<const int4*>(&B.mtid[lane * WT_ITEMS]);
            const int4 c = *reinterpret_cast<const int4*>(&B.qlen[lane * WT_ITEMS]);
            tid[0] = a.x; tid[1] = a.y; tid[2] = a.z; tid[3] = a.w;
            mtid[0] = b.x; mtid[1] = b.y; mtid[2] = b.z; mtid[3] = b.w;
            if (PACKED) {
                const u32 w[WT_ITEMS] = {(u32)c.x, (u32)c.y, (u32)c.z, (u32)c.w};
#pragma unroll
                for (int i = 0; i < WT_ITEMS; ++i) { flag[i] = w[i] & 0xfffu; mapq[i] = (w[i] >> 12) & 0xffu; qlen[i] = (int)(w[i] >> 20); }
            } else {
                const uint2 f = *reinterpret_cast<const uint2*>(&B.flag[lane * WT_ITEMS]);
                const u32 m = *reinterpret_cast<const u32*>(&B.mapq[lane * WT_ITEMS]);
                qlen[0] = c.x; qlen[1] = c.y; qlen[2] = c.z; qlen[3] = c.w;
                flag[0] = f.x & 0xffffu; flag[1] = f.x >> 16; flag[2] = f.y & 0xffffu; flag[3] = f.y >> 16;
                mapq[0] = m & 0xffu; mapq[1] = (m >> 8) & 0xffu; mapq[2] = (m >> 16) & 0xffu; mapq[3] = m >> 24;
            }
        }

        // ---- per-record classification (CreateGraph.py:118-206); only word 0 of a contig row is needed:
        // 0 <=> absent (state ABSENT == 0) or out of range (:118-130).  Kept branchy on purpose: the
        // same-contig half of the records leaves early (a predicated version issued 24 % more instructions)
        const int t0 = tid[0];
        const int xa = ((u32)t0 < n_contigs) ? __ldg(reinterpret_cast<const int*>(P.rows + t0)) : 0;
        u32 elig = 0, cov = 0;   // one bit per item
#pragma unroll
        for (int i = 0; i < WT_ITEMS; ++i) {
            int x1 = xa;
            if (i > 0 && tid[i] != t0) x1 = ((u32)tid[i] < n_contigs) ? __ldg(reinterpret_cast<const int*>(P.rows + tid[i])) : 0;
            const int mq = (int)mapq[i];
            const bool covered = (mq >= P.min_mapq) || mq == 0;                                           // :138
            if (tid[i] == mtid[i]) {   // same contig: same scaffold, no link, no fishy pair
                if (row_state(x1) != BESST_CTG_ABSENT) { c_valid++; if (covered) cov |= 1u << i; }
                continue;
            }
            const int x2 = ((u32)mtid[i] < n_contigs) ? __ldg(reinterpret_cast<const int*>(P.rows + mtid[i])) : 0;
            if (row_state(x1) == BESST_CTG_ABSENT || row_state(x2) == BESST_CTG_ABSENT) continue;          // :127-130
            c_valid++;
            if (covered) cov |= 1u << i;
            const u32 f = flag[i];
            const bool unmapped = f & 0x4u;
            const bool diff_scaf = ((u32)(x1 ^ x2) >> 3) != 0;
            if (unmapped && (f & 0x40u) && diff_scaf) {                                                   // :141-163
                const u32 n1 = 2u * (u32)row_scaf(x1) + (u32)side_only(row_dir(x1), !(f & 0x10u), P.orientation);
                const u32 n2 = 2u * (u32)row_scaf(x2) + (u32)side_only(row_dir(x2), !(f & 0x20u), P.orientation);
                const u64 key = n1 < n2 ? (((u64)n1 << 32) | n2) : (((u64)n2 << 32) | n1);
                const u64 slot = atomicAdd(&P.globals[1], 1ull);
                if ((long long)slot < P.fishy_cap) P.fishy[slot] = key;
                c_fishy++;
            }
            if (mq == 0) c_nonuniq++;                                                                     // :166-167
            if ((f & 0x80u) && !unmapped && mq >= P.min_mapq) {                                           // :169
                // :170 large-large needs different scaffolds; :184-206 (extend_paths) small-small needs different
                // scaffolds, exactly one small always qualifies.  s1|s2: 1 both large, 2 both small, 3 one of each
                const u32 so = ((u32)x1 | (u32)x2) & 3u;
                const bool e = so == 3u ? P.extend != 0 : (diff_scaf && (so == 1u || P.extend != 0));
                if (e) elig |= 1u << i;
            }
        }

        // ---- coverage: warp-aggregated 64-bit atomics (:138-139) ----------------------------
        {
            int s0 = 0;
            u32 rest = 0;
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i) {
                const bool c = cov >> i & 1u, own = tid[i] == t0;
                s0 += (c && own) ? qlen[i] : 0;
                rest |= ((c && !own) ? 1u : 0u) << i;
            }
            // common case: the whole tile lies on one contig -> one reduction, one atomic
            const int tw = __shfl_sync(0xffffffffu, t0, 0);
            if (__all_sync(0xffffffffu, s0 == 0 || t0 == tw)) {
                const int sum = __reduce_add_sync(0xffffffffu, s0);
                if (lane == 0 && sum != 0) atomicAdd(&P.aligned[tw], (u64)(long long)sum);
            } else {
                const u32 act = __ballot_sync(0xffffffffu, s0 != 0);
                if (s0 != 0) {
                    const u32 peers = __match_any_sync(act, t0);
                    const int sum = __reduce_add_sync(peers, s0);
                    if (lane == __ffs(peers) - 1) atomicAdd(&P.aligned[t0], (u64)(long long)sum);
                }
            }
            if (__any_sync(0xffffffffu, rest != 0)) {   // a lane's 4 records straddle contigs: rare
#pragma unroll
                for (int i = 1; i < WT_ITEMS; ++i)
                    if (rest >> i & 1u) atomicAdd(&P.aligned[tid[i]], (u64)(long long)qlen[i]);
            }
        }

        // ---- compact the CreateEdge candidates of the tile, BAM order, one per lane ------------
        const int my_c = __popc(elig);
        int incl_c = my_c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl_c, off);
            if (lane >= off) incl_c += t;
        }
        const int total_c = __shfl_sync(0xffffffffu, incl_c, 31);
        int n_out = 0;
        u32 first_flags = 0;
        int first_o1 = 0, first_o2 = 0, carry_o1 = 0, carry_o2 = 0;
        if (total_c > 0) {
            int slot = incl_c - my_c;
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i)
                if (elig >> i & 1u) W.cand[slot++] = (unsigned char)(lane * WT_ITEMS + i);
            if (can_demand && !pos_ready) {   // pos/mpos were not fetched ahead for this tile
                if (elect_one()) {
                    const long long r0 = wt * WT;
                    mbar_expect_tx(bar0 + 16u, TILE_BYTES_POS);
                    bulk_g2s(smem_u32(&B.pos[0]), P.rec.pos + r0, 4 * WT, bar0 + 16u);
                    bulk_g2s(smem_u32(&B.mpos[0]), P.rec.mpos + r0, 4 * WT, bar0 + 16u);
                }
                mbar_wait(bar0 + 16u, demand_par);
                demand_par ^= 1u;
                ++c_postiles;
            }
            __syncwarp();

            // ---- CreateEdge, one call per lane: observations (:816-833), duplicate test against the
            // previous call (:835-838), acceptance test (:840), counters.  The tile's first call has no
            // in-tile predecessor: assumed "not a duplicate" here and settled by the aggregate scan.
            bool have_carry = false;
            int4* const out = reinterpret_cast<int4*>(P.scratch) + wt * WT;
            for (int r = 0; r < total_c; r += 32) {
                const int k = r + lane;
                const bool active = k < total_c;
                int o1 = 0, o2 = 0;
                u32 nu = 0, nv = 0;
                bool mq0 = false, both_large = false;
                if (active) {
                    const int j = W.cand[k];
                    const u32 pw = PACKED ? (u32)B.qlen[j] : 0u;
                    const u32 fl = PACKED ? (pw & 0xfffu) : (u32)B.flag[j];
                    mq0 = PACKED ? (((pw >> 12) & 0xffu) == 0u) : (B.mapq[j] == 0);
                    const int4 r1 = __ldg(P.rows + B.tid[j]);   // L1 hits: word 0 was gathered a moment ago
                    const int4 r2 = __ldg(P.rows + B.mtid[j]);
                    both_large = (((u32)r1.x | (u32)r2.x) & 3u) == (u32)BESST_CTG_LARGE;   // neither is absent here
                    int s1, s2;
                    pos_dir<INT_RL>(row_dir(r1.x), !(fl & 0x10u), P.orientation, r1.y, B.pos[j], r1.w, r1.z, P.read_len, P.read_len_i, o1, s1);
                    pos_dir<INT_RL>(row_dir(r2.x), !(fl & 0x20u), P.orientation, r2.y, B.mpos[j], r2.w, r2.z, P.read_len, P.read_len_i, o2, s2);
                    nu = 2u * (u32)row_scaf(r1.x) + (u32)s1;
                    nv = 2u * (u32)row_scaf(r2.x) + (u32)s2;
                }
                int p1 = __shfl_up_sync(0xffffffffu, o1, 1), p2 = __shfl_up_sync(0xffffffffu, o2, 1);
                bool has_prev = true;
                if (lane == 0) { p1 = carry_o1; p2 = carry_o2; has_prev = have_carry; }
                bool accepted = false;
                if (active) {
                    c_calls++;
                    c_nonuniq_scaf += mq0 ? 1 : 0;
                    const bool dup = has_prev && o1 == p1 && o2 == p2;
                    c_dups += dup ? 1 : 0;
                    const bool is_dupl = dup && P.detect_dup;
                    const bool pass = o1 > 25 && o2 > 25 && (thr_all || (u32)o1 + (u32)o2 < thr_u);
                    const bool second = both_large && second_on;
                    if (!has_prev) {   // the tile's first call
                        first_flags = AGG_HAS | (pass ? AGG_PASS : 0u) | (second ? AGG_SECOND : 0u) | (mq0 ? AGG_MQ0 : 0u);
                        first_o1 = o1; first_o2 = o2;
                    }
                    if (!is_dupl) {
                        // the second call into G_prime (:180-183) sees prev_obs reset to -1: a duplicate only of (-1,-1)
                        const bool dup2 = second && o1 == -1 && o2 == -1;
                        const int calls = 1 + ((second && !(dup2 && P.detect_dup)) ? 1 : 0);
                        c_nonuniq_scaf += (second && mq0) ? 1 : 0;
                        c_dups += dup2 ? 1 : 0;
                        c_count += pass ? calls : 0;
                        c_toolong += pass ? 0 : calls;
                        accepted = pass;
                    }
                }
                const u32 bal = __ballot_sync(0xffffffffu, accepted);
                if (accepted) {
                    int4 t;
                    const bool fw = nu < nv;
                    t.x = (int)(fw ? nu : nv); t.y = (int)(fw ? nv : nu); t.z = fw ? o1 : o2; t.w = fw ? o2 : o1;
                    out[n_out + __popc(bal & lt_mask)] = t;
                }
                n_out += __popc(bal);
                const int last_lane = (total_c - r > 32) ? 31 : (total_c - r - 1);
                carry_o1 = __shfl_sync(0xffffffffu, o1, last_lane);
                carry_o2 = __shfl_sync(0xffffffffu, o2, last_lane);
                have_carry = true;
            }
        }
        __syncwarp();   // the stage and the candidate list are reused two / one tiles from now

        // ---- aggregate (the first call of the tile sits in lane 0 of the first round) -----------------
        if (lane == 0) {
            int4* a = reinterpret_cast<int4*>(P.aggs + wt);
            a[0] = make_int4((int)first_flags, first_o1, first_o2, carry_o1);
            a[1] = make_int4(carry_o2, n_out, 0, 0);
            a[2] = make_int4(0, 0, 0, 0);
        }
        return total_c > 0;
    };

    // ---- full tiles: TMA-staged, double buffered, handed out dynamically in batches of K1T_BATCH
    // consecutive tiles (one ticket per warp and batch; a ticket per tile would serialise 3 M same-address
    // atomics) after a static first batch: warps progress at different speeds and a static split leaves
    // a quarter of the warp slots idle towards the end.  The next batch's ticket is requested a batch ahead.
    const long long n_full = vec_ok ? n / WT : 0;
    auto issue = [&](const long long tile, const int s, const bool want_pos) {
        c_postiles += want_pos ? 1 : 0;
        if (elect_one()) {
            const long long r0 = tile * WT;
            const u32 bar = bar0 + 8u * s;
            const u32 dst = smem_u32(&W.buf[s]);
            mbar_expect_tx(bar, (PACKED ? 3u * 4u * WT : TILE_BYTES_NOPOS) + (want_pos ? TILE_BYTES_POS : 0u));
            bulk_g2s(dst + 0 * 4 * WT, P.rec.tid + r0, 4 * WT, bar);
            bulk_g2s(dst + 1 * 4 * WT, P.rec.mtid + r0, 4 * WT, bar);
            if (PACKED) bulk_g2s(dst + 2 * 4 * WT, P.rec.packed + r0, 4 * WT, bar);
            else bulk_g2s(dst + 2 * 4 * WT, P.rec.qlen + r0, 4 * WT, bar);
            if (want_pos) {
                bulk_g2s(dst + 3 * 4 * WT, P.rec.pos + r0, 4 * WT, bar);
                bulk_g2s(dst + 4 * 4 * WT, P.rec.mpos + r0, 4 * WT, bar);
            }
            if (!PACKED) {
                bulk_g2s(dst + 5 * 4 * WT, P.rec.flag + r0, 2 * WT, bar);
                bulk_g2s(dst + 5 * 4 * WT + 2 * WT, P.rec.mapq + r0, WT, bar);
            }
        }
    };
    u32* const ticket = reinterpret_cast<u32*>(P.globals + 2);
    u32 cand_hist = 7u;   // bit t: the t-th last tile had a CreateEdge candidate
    u32 st_pos = 1u;      // bit s: stage s has (or is getting) its pos/mpos columns
    long long wt = (long long)warp_global * K1T_BATCH;
    int b_left = K1T_BATCH - 1;   // tiles left in the current batch after wt
    u32 t_raw = 0;
    if (lane == 0) t_raw = atomicAdd(ticket, 1u);
    if (wt < n_full) issue(wt, 0, true);
    for (u32 it = 0; wt < n_full; ++it) {
        const int s = (int)(it & 1u);
        long long nx = wt + 1;
        if (b_left == 0) {   // next batch: (ticket + number of static batches) * batch size
            nx = ((long long)n_warps + (long long)__shfl_sync(0xffffffffu, t_raw, 0)) * K1T_BATCH;
            b_left = K1T_BATCH;
            if (lane == 0) t_raw = atomicAdd(ticket, 1u);
        }
        --b_left;
        if (nx < n_full) {
            const bool want = (cand_hist & 7u) != 0;
            issue(nx, s ^ 1, want);
            st_pos = (st_pos & ~(2u >> s)) | ((want ? 2u : 0u) >> s);   // bit s ^ 1
        }
        mbar_wait(bar0 + 8u * s, (it >> 1) & 1u);
        const bool had = process(wt, W.buf[s], (st_pos >> s) & 1u, true);
        cand_hist = (cand_hist << 1) | (had ? 1u : 0u);
        wt = nx;
    }
    // ---- leftovers: the ragged last tile, or every tile when a column is not 16-byte aligned: plain
    // bounds-checked loads into stage 0 (no bulk copy is outstanding any more)
    for (long long t = n_full + warp_global; t < P.n_tiles; t += n_warps) {
        TileBuf& B = W.buf[0];
        const long long idx0 = t * WT + (long long)lane * WT_ITEMS;
#pragma unroll
        for (int i = 0; i < WT_ITEMS; ++i) {
            const long long r = idx0 + i;
            const bool in = r < n;
            const int q = lane * WT_ITEMS + i;
            B.tid[q] = in ? __ldg(P.rec.tid + r) : -1;
            B.mtid[q] = in ? __ldg(P.rec.mtid + r) : -1;
            B.pos[q] = in ? __ldg(P.rec.pos + r) : 0;
            B.mpos[q] = in ? __ldg(P.rec.mpos + r) : 0;
            if (PACKED) {
                B.qlen[q] = in ? (int)__ldg(P.rec.packed + r) : 0;
            } else {
                B.qlen[q] = in ? __ldg(P.rec.qlen + r) : 0;
                B.flag[q] = in ? __ldg(P.rec.flag + r) : (unsigned short)0;
                B.mapq[q] = in ? __ldg(P.rec.mapq + r) : (unsigned char)0;
            }
        }
        __syncwarp();
        ++c_postiles;
        process(t, B, true, false);
    }
    if (lane == 0 && c_postiles) atomicAdd(&P.counters[BESST_CNT_POS_TILES], (u64)c_postiles);

    // ---- flush the per-thread counters ------------------------------------------------------------
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

// ---- scan over aggregates ------------------------------------------------------------------------
// One block of SC_THREADS threads scans SC_THREADS aggregates: thread t owns aggregate t.  `carry`
// is the last CreateEdge observation before the block's first aggregate (has = 0: unknown, the
// block's first call stays unresolved and is described in the block's own aggregate).
struct ScanResult {
    u32 cnt_adj;      // accepted tuples of this aggregate after the boundary verdict
    u64 excl;         // sum of cnt_adj over the preceding aggregates of the block
    bool drop_first;  // boundary duplicate that was an accepted tuple: skip it
    Last before;      // last CreateEdge call before this aggregate (has = 0: unknown)
    Agg block;        // aggregate of the whole block (valid in every thread)
};

struct ScanSmem {
    Last w_last[SC_THREADS / 32];
    u32 w_cnt[SC_THREADS / 32];
    int w_corr[4][SC_THREADS / 32];
    int first_thread;
    Agg first;
};

__device__ __forceinline__ ScanResult block_scan_aggs(const Agg& a, Last carry, int detect_dup, ScanSmem& S) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool has = a.flags & AGG_HAS;
    Last mine = {has ? 1 : 0, a.last_o1, a.last_o2};
    const Last incl = warp_scan_last(mine, lane);
    Last excl;
    excl.has = __shfl_up_sync(0xffffffffu, incl.has, 1);
    excl.o1 = __shfl_up_sync(0xffffffffu, incl.o1, 1);
    excl.o2 = __shfl_up_sync(0xffffffffu, incl.o2, 1);
    if (lane == 0) excl.has = 0;
    if (lane == 31) S.w_last[warp] = incl;
    if (t == 0) S.first_thread = -1;
    __syncthreads();
    Last before = carry;
    for (int w = 0; w < warp; ++w)
        if (S.w_last[w].has) before = S.w_last[w];
    if (!excl.has) excl = before;

    // boundary verdict for this aggregate's first call
    int dups = a.dups, count_corr = a.count_corr, toolong_corr = a.toolong_corr, nus_corr = a.nus_corr;
    u32 cnt_adj = a.cnt;
    bool drop_first = false;
    if (has) {
        if (excl.has) {
            if (a.first_o1 == excl.o1 && a.first_o2 == excl.o2) {
                dups += 1;
                if (detect_dup) {   // the call returned early (:835-838): undo what the tile assumed
                    const int calls = 1 + ((a.flags & AGG_SECOND) ? 1 : 0);
                    if (a.flags & AGG_PASS) { count_corr -= calls; cnt_adj -= 1; drop_first = true; }
                    else toolong_corr -= calls;
                    if ((a.flags & AGG_SECOND) && (a.flags & AGG_MQ0)) nus_corr -= 1;
                }
            }
        } else {
            S.first_thread = t;   // unique: the first aggregate with a call and nothing known before it
        }
    }
    // sums
    u32 incl_cnt = cnt_adj;
    int c0 = dups, c1 = count_corr, c2 = toolong_corr, c3 = nus_corr;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u32 v = __shfl_up_sync(0xffffffffu, incl_cnt, off);
        if (lane >= off) incl_cnt += v;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, off);
        c1 += __shfl_xor_sync(0xffffffffu, c1, off);
        c2 += __shfl_xor_sync(0xffffffffu, c2, off);
        c3 += __shfl_xor_sync(0xffffffffu, c3, off);
    }
    if (lane == 31) S.w_cnt[warp] = incl_cnt;
    if (lane == 0) { S.w_corr[0][warp] = c0; S.w_corr[1][warp] = c1; S.w_corr[2][warp] = c2; S.w_corr[3][warp] = c3; }
    __syncthreads();
    if (S.first_thread == t) S.first = a;
    __syncthreads();
    ScanResult R;
    u64 pre = 0, tot = 0;
    int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    Last blk_last = {0, 0, 0};
    for (int w = 0; w < SC_THREADS / 32; ++w) {
        if (w < warp) pre += S.w_cnt[w];
        tot += S.w_cnt[w];
        s0 += S.w_corr[0][w]; s1 += S.w_corr[1][w]; s2 += S.w_corr[2][w]; s3 += S.w_corr[3][w];
        if (S.w_last[w].has) blk_last = S.w_last[w];
    }
    R.cnt_adj = cnt_adj;
    R.excl = pre + incl_cnt - cnt_adj;
    R.drop_first = drop_first;
    R.before = excl;
    R.block.flags = 0;
    R.block.first_o1 = R.block.first_o2 = 0;
    if (S.first_thread >= 0) {
        R.block.flags = S.first.flags & (AGG_PASS | AGG_SECOND | AGG_MQ0);
        R.block.first_o1 = S.first.first_o1;
        R.block.first_o2 = S.first.first_o2;
    }
    if (blk_last.has) R.block.flags |= AGG_HAS;
    // a block whose calls were all resolved against `carry` has no open first call; a block scanned
    // without carry has one exactly when it has any call
    R.block.last_o1 = blk_last.o1; R.block.last_o2 = blk_last.o2;
    R.block.cnt = (u32)tot;
    R.block.dups = s0; R.block.count_corr = s1; R.block.toolong_corr = s2; R.block.nus_corr = s3;
    R.block.pad0 = R.block.pad1 = 0;
    return R;
}

__device__ __forceinline__ Agg load_agg(const Agg* p, long long i, long long n) {
    Agg a;
    if (i < n) {
        const int4* q = reinterpret_cast<const int4*>(p + i);
        const int4 x = q[0], y = q[1], z = q[2];
        a.flags = (u32)x.x; a.first_o1 = x.y; a.first_o2 = x.z; a.last_o1 = x.w;
        a.last_o2 = y.x; a.cnt = (u32)y.y; a.dups = y.z; a.count_corr = y.w;
        a.toolong_corr = z.x; a.nus_corr = z.y; a.pad0 = a.pad1 = 0;
    } else {
        a.flags = 0; a.first_o1 = a.first_o2 = a.last_o1 = a.last_o2 = 0; a.cnt = 0;
        a.dups = a.count_corr = a.toolong_corr = a.nus_corr = 0; a.pad0 = a.pad1 = 0;
    }
    return a;
}

// level 0 -> level 1: one aggregate per chunk of SC_THREADS tiles
__global__ void __launch_bounds__(SC_THREADS) k_tile_reduce(const Agg* __restrict__ aggs, long long n_tiles, Agg* chunk_aggs,
                                                            int detect_dup) {
    __shared__ ScanSmem S;
    const long long i = (long long)blockIdx.x * SC_THREADS + threadIdx.x;
    const Agg a = load_agg(aggs, i, n_tiles);
    const Last none = {0, 0, 0};
    const ScanResult R = block_scan_aggs(a, none, detect_dup, S);
    if (threadIdx.x == 0) chunk_aggs[blockIdx.x] = R.block;
}

// level 1: one block walks the chunk aggregates in groups of SC_THREADS, carrying (last call,
// offset, corrections); writes what every chunk receives, the totals and the next rank's halo
__global__ void __launch_bounds__(SC_THREADS) k_chunk_resolve(const Agg* __restrict__ chunk_aggs, long long n_chunks,
                                                              ChunkIn* chunk_in, u64* tile_off_end, int halo1, int halo2,
                                                              int detect_dup, u64* counters, u64* globals) {
    __shared__ ScanSmem S;
    __shared__ Last s_carry;
    __shared__ u64 s_off;
    __shared__ long long s_corr[4];
    __shared__ unsigned long long s_first_chunk;   // first chunk holding a CreateEdge call
    if (threadIdx.x == 0) {
        s_first_chunk = ~0ull;
        s_carry.has = 1; s_carry.o1 = halo1; s_carry.o2 = halo2;   // counters(..., prev_obs1=-1, prev_obs2=-1) :98
        s_off = 0;
        s_corr[0] = s_corr[1] = s_corr[2] = s_corr[3] = 0;
    }
    __syncthreads();
    for (long long base = 0; base < n_chunks; base += SC_THREADS) {
        const long long i = base + threadIdx.x;
        const Agg a = load_agg(chunk_aggs, i, n_chunks);
        if (a.flags & AGG_HAS) atomicMin(&s_first_chunk, (unsigned long long)i);
        const Last carry = s_carry;
        const u64 off0 = s_off;
        const ScanResult R = block_scan_aggs(a, carry, detect_dup, S);
        if (i < n_chunks) {
            ChunkIn in;
            in.carry_o1 = R.before.o1; in.carry_o2 = R.before.o2; in.offset = off0 + R.excl;
            chunk_in[i] = in;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (R.block.flags & AGG_HAS) { s_carry.o1 = R.block.last_o1; s_carry.o2 = R.block.last_o2; }
            s_off = off0 + R.block.cnt;
            s_corr[0] += R.block.dups; s_corr[1] += R.block.count_corr; s_corr[2] += R.block.toolong_corr;
            s_corr[3] += R.block.nus_corr;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        globals[0] = s_off;
        *tile_off_end = s_off;
        counters[BESST_CNT_DUPLICATES] += (u64)s_corr[0];
        counters[BESST_CNT_COUNT] += (u64)s_corr[1];
        counters[BESST_CNT_TOO_LONG] += (u64)s_corr[2];
        counters[BESST_CNT_NON_UNIQUE_SCAF] += (u64)s_corr[3];
        counters[BESST_CNT_LAST_OBS1] = (u64)(long long)s_carry.o1;
        counters[BESST_CNT_LAST_OBS2] = (u64)(long long)s_carry.o2;
        long long f1 = 0, f2 = 0;
        if (s_first_chunk != ~0ull) { f1 = chunk_aggs[s_first_chunk].first_o1; f2 = chunk_aggs[s_first_chunk].first_o2; }
        counters[BESST_CNT_FIRST_OBS1] = (u64)f1;
        counters[BESST_CNT_FIRST_OBS2] = (u64)f2;
    }
}

// level 0 again, now with what each chunk receives: exact output offset of every tile
// also: block_tile0[g] = the tile holding the tuple with ordinal 2048 * g (where the grouping block g of the
// run-merge bucket starts reading the scratch runs)
__global__ void __launch_bounds__(SC_THREADS) k_tile_offsets(const Agg* __restrict__ aggs, long long n_tiles,
                                                             const ChunkIn* __restrict__ chunk_in, u64* tile_off,
                                                             int detect_dup, u32* block_tile0) {
    __shared__ ScanSmem S;
    const long long i = (long long)blockIdx.x * SC_THREADS + threadIdx.x;
    const Agg a = load_agg(aggs, i, n_tiles);
    const ChunkIn in = chunk_in[blockIdx.x];
    const Last carry = {1, in.carry_o1, in.carry_o2};
    const ScanResult R = block_scan_aggs(a, carry, detect_dup, S);
    if (i < n_tiles) {
        const u64 a = in.offset + R.excl;
        tile_off[i] = a | (R.drop_first ? OFF_DROP : 0ull);
        const u64 g = (a + 2047) >> 11;   // a tile keeps at most 128 tuples: it can hold at most one block start
        if ((g << 11) < a + R.cnt_adj) block_tile0[g] = (u32)i;
    }
}

// tile-local scratch runs -> the final BAM-ordered tuple array
__global__ void __launch_bounds__(256) k_compact_tuples(const besst_link_tuple* __restrict__ scratch,
                                                        const u64* __restrict__ tile_off, long long n_tiles,
                                                        besst_link_tuple* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int4* src_base = reinterpret_cast<const int4*>(scratch);
    int4* dst_base = reinterpret_cast<int4*>(out);
    for (long long wt = warp_global; wt < n_tiles; wt += n_warps) {
        const u64 a = __ldg(tile_off + wt), b = __ldg(tile_off + wt + 1);
        const long long off = (long long)(a & OFF_MASK);
        const int cnt = (int)((long long)(b & OFF_MASK) - off);
        const int4* src = src_base + wt * WT + ((a & OFF_DROP) ? 1 : 0);
        for (int j = lane; j < cnt; j += 32) dst_base[off + j] = __ldg(src + j);
    }
}

// ---- multi-GPU: stable partition of tuples / fishy keys by destination rank ----------------------
// dest = hash(u, v) mod world.  Order inside a destination bucket is BAM order, so concatenating the
// buckets received from ranks 0..W-1 reproduces the global BAM order per edge.
constexpr int PT_THREADS = 256;
constexpr int PT_ITEMS = 8;
constexpr int PT_TILE = PT_THREADS * PT_ITEMS;
constexpr int PT_MAX_WORLD = 16;

__device__ __forceinline__ u32 edge_dest(u32 u, u32 v, int world) { return besst_edge_dest(u, v, world); }

template <bool TUPLES>
__device__ __forceinline__ u32 item_dest(const void* in, long long i, int world) {
    if (TUPLES) {
        const uint2 uv = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const besst_link_tuple*>(in) + i));
        return edge_dest(uv.x, uv.y, world);
    }
    const u64 k = __ldg(reinterpret_cast<const u64*>(in) + i);
    return edge_dest((u32)(k >> 32), (u32)k, world);
}

// per-tile destination histogram: counts[d * n_tiles + tile]
template <bool TUPLES>
__global__ void __launch_bounds__(PT_THREADS) k_partition_count(const void* __restrict__ in, long long n, int world,
                                                                u32* __restrict__ counts, int n_tiles) {
    __shared__ u32 s_h[PT_MAX_WORLD];
    if (threadIdx.x < PT_MAX_WORLD) s_h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * PT_TILE;
    u32 local[PT_MAX_WORLD];
#pragma unroll
    for (int d = 0; d < PT_MAX_WORLD; ++d) local[d] = 0;
#pragma unroll
    for (int k = 0; k < PT_ITEMS; ++k) {
        const long long i = base + k * PT_THREADS + threadIdx.x;
        if (i < n) {
            const u32 d = item_dest<TUPLES>(in, i, world);
#pragma unroll
            for (int q = 0; q < PT_MAX_WORLD; ++q) local[q] += (q == (int)d);
        }
    }
#pragma unroll
    for (int d = 0; d < PT_MAX_WORLD; ++d) {
        u32 v = local[d];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_h[d], v);
    }
    __syncthreads();
    if (threadIdx.x < world) counts[(size_t)threadIdx.x * n_tiles + blockIdx.x] = s_h[threadIdx.x];
}

// exclusive scan of counts[0 .. len) in place (64-bit carry, single block); bucket totals -> totals[d]
__global__ void __launch_bounds__(1024) k_partition_scan(u32* counts, int n_tiles, int world, u64* starts, u64* totals) {
    __shared__ u64 s_w[32];
    __shared__ u64 s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int d = 0; d < world; ++d) {
        const u64 bucket_start = s_carry;
        if (threadIdx.x == 0) starts[d] = bucket_start;
        u32* c = counts + (size_t)d * n_tiles;
        for (int base = 0; base < n_tiles; base += 1024) {
            const int i = base + threadIdx.x;
            const u64 v = i < n_tiles ? c[i] : 0;
            u64 incl = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const u64 t = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += t;
            }
            if (lane == 31) s_w[warp] = incl;
            __syncthreads();
            u64 pre = 0, tot = 0;
            for (int w = 0; w < 32; ++w) { if (w < warp) pre += s_w[w]; tot += s_w[w]; }
            const u64 carry = s_carry;
            // offsets inside a bucket fit 32 bits (n < 2^30); the bucket start is added by the scatter
            if (i < n_tiles) c[i] = (u32)(carry - bucket_start + pre + incl - v);
            __syncthreads();
            if (threadIdx.x == 0) s_carry = carry + tot;
            __syncthreads();
        }
        if (threadIdx.x == 0) totals[d] = s_carry - bucket_start;
    }
}

template <bool TUPLES>
__global__ void __launch_bounds__(PT_THREADS) k_partition_scatter(const void* __restrict__ in, long long n, int world,
                                                                  const u32* __restrict__ offs, int n_tiles,
                                                                  const u64* __restrict__ starts, void* __restrict__ out,
                                                                  u32* __restrict__ ord_out) {
    __shared__ u32 s_wcnt[PT_THREADS / 32][PT_MAX_WORLD];
    __shared__ u64 s_base[PT_MAX_WORLD];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * PT_TILE + (long long)threadIdx.x * PT_ITEMS;   // blocked: keeps order
    u32 dest[PT_ITEMS];
    u32 local[PT_MAX_WORLD];
#pragma unroll
    for (int d = 0; d < PT_MAX_WORLD; ++d) local[d] = 0;
#pragma unroll
    for (int k = 0; k < PT_ITEMS; ++k) {
        dest[k] = 0xffffffffu;
        if (base + k < n) {
            dest[k] = item_dest<TUPLES>(in, base + k, world);
#pragma unroll
            for (int q = 0; q < PT_MAX_WORLD; ++q) local[q] += (q == (int)dest[k]);
        }
    }
    // exclusive scan over threads, per destination
    u32 excl[PT_MAX_WORLD];
#pragma unroll
    for (int d = 0; d < PT_MAX_WORLD; ++d) {
        u32 incl = local[d];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        excl[d] = incl - local[d];
        if (lane == 31) s_wcnt[warp][d] = incl;
    }
    if (threadIdx.x < world) s_base[threadIdx.x] = starts[threadIdx.x] + offs[(size_t)threadIdx.x * n_tiles + blockIdx.x];
    __syncthreads();
#pragma unroll
    for (int d = 0; d < PT_MAX_WORLD; ++d)
        for (int w = 0; w < warp; ++w) excl[d] += s_wcnt[w][d];
#pragma unroll
    for (int k = 0; k < PT_ITEMS; ++k) {
        if (dest[k] == 0xffffffffu) continue;
        u32 r = 0;
#pragma unroll
        for (int q = 0; q < PT_MAX_WORLD; ++q)
            if (q == (int)dest[k]) { r = excl[q]; excl[q] += 1; }
        const u64 o = s_base[dest[k]] + r;
        if (TUPLES) {
            reinterpret_cast<int4*>(out)[o] = __ldg(reinterpret_cast<const int4*>(in) + base + k);
            if (ord_out) ord_out[o] = (u32)(base + k);   // ordinal in this rank's BAM-ordered tuple stream
        } else {
            reinterpret_cast<u64*>(out)[o] = __ldg(reinterpret_cast<const u64*>(in) + base + k);
        }
    }
}

template <bool TUPLES>
int partition_impl(besst_ctx* ctx, const void* in, int64_t n, int world, void* out, u32* ord_out, int64_t* counts_host) {
    if (counts_host) for (int d = 0; d < world; ++d) counts_host[d] = 0;
    if (n == 0) return BESST_OK;
    const int n_tiles = (int)((n + PT_TILE - 1) / PT_TILE);
    BESST_CUDA_TRY(ctx, ctx->part_state.ensure(4 * (size_t)n_tiles * world + 16 * PT_MAX_WORLD + 64));
    u64* starts = ctx->part_state.as<u64>();
    u64* totals = starts + PT_MAX_WORLD;
    u32* counts = reinterpret_cast<u32*>(totals + PT_MAX_WORLD);
    { KTimer kt(ctx, BESST_K_PARTITION); k_partition_count<TUPLES><<<n_tiles, PT_THREADS, 0, ctx->stream>>>(in, n, world, counts, n_tiles); }
    { KTimer kt(ctx, BESST_K_PARTITION); k_partition_scan<<<1, 1024, 0, ctx->stream>>>(counts, n_tiles, world, starts, totals); }
    { KTimer kt(ctx, BESST_K_PARTITION); k_partition_scatter<TUPLES><<<n_tiles, PT_THREADS, 0, ctx->stream>>>(in, n, world, counts, n_tiles, starts, out, ord_out); }
    BESST_CUDA_TRY(ctx, cudaGetLastError());
    if (!counts_host) return BESST_OK;   // queued only: the totals stay in part_state[PT_MAX_WORLD ..] for the caller's one read
    u64 h[PT_MAX_WORLD];
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(h, totals, 8 * world, cudaMemcpyDeviceToHost, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int d = 0; d < world; ++d) counts_host[d] = (int64_t)h[d];
    return BESST_OK;
}

}  // namespace

// the fishy keys only, queued without a host read; *totals_device: u64[world] bucket sizes (valid once the stream gets there)
int besst_launch_partition_fishy_async(besst_ctx* ctx, int world, uint64_t* out_fishy, const uint64_t** totals_device) {
    *totals_device = nullptr;
    if (ctx->n_fishy_keys == 0) return BESST_OK;
    const int rc = partition_impl<false>(ctx, ctx->fishy_keys.p, ctx->n_fishy_keys, world, out_fishy, nullptr, nullptr);
    if (rc) return rc;
    *totals_device = reinterpret_cast<const uint64_t*>(ctx->part_state.as<u64>() + PT_MAX_WORLD);
    return BESST_OK;
}

int besst_launch_partition(besst_ctx* ctx, int world, besst_link_tuple* out_tuples, uint32_t* out_ordinals,
                           uint64_t* out_fishy, int64_t* tuple_counts, int64_t* fishy_counts) {
    if (world < 1 || world > PT_MAX_WORLD) { ctx->err = "partition: world size must be 1..16"; return BESST_E_INVALID; }
    if (ctx->n_tuples >= (1ll << 30) || ctx->n_fishy_keys >= (1ll << 30)) { ctx->err = "partition: more than 2^30 items"; return BESST_E_INVALID; }
    int rc = BESST_OK;
    if (out_tuples) rc = besst_ensure_tuples(ctx);
    if (rc) return rc;
    if (out_tuples) rc = partition_impl<true>(ctx, ctx->tuples.p, ctx->n_tuples, world, out_tuples, out_ordinals, tuple_counts);
    else for (int d = 0; d < world; ++d) tuple_counts[d] = 0;   // run-level exchange: only the fishy keys travel as keys
    if (rc) return rc;
    return partition_impl<false>(ctx, ctx->fishy_keys.p, ctx->n_fishy_keys, world, out_fishy, nullptr, fishy_counts);
}

// ---- launcher: begin / slice* / finish --------------------------------------------------------------
// The record range can be fed in slices (besst_extract_slice, r0 a multiple of the tile size) so
// that a host-buffer call overlaps the H2D copy of slice k+1 with K1 on slice k.
static int64_t tiles_of(int64_t n) { return (n + WT - 1) / WT; }

int besst_extract_begin(besst_ctx* ctx, const besst_lib_params& p, int64_t n) {
    (void)p;
    const int64_t n_tiles = tiles_of(n);
    const int64_t n_chunks = (n_tiles + SC_THREADS - 1) / SC_THREADS;
    if (n_chunks > 0x7fffffff) { ctx->err = "too many records for one call"; return BESST_E_INVALID; }
    // coverage and counters share ONE allocation (aligned_len[C + 1] followed by the counters): one memset here, and a
    // multi-GPU driver reduces both with a single all-reduce over the contiguous span
    const size_t c1 = (size_t)(ctx->n_contigs + 1);
    BESST_CUDA_TRY(ctx, ctx->aligned.ensure(sizeof(u64) * (c1 + BESST_N_COUNTERS + 8)));
    ctx->counters.p = ctx->aligned.as<u64>() + c1;   // a view: never freed on its own
    ctx->counters.cap = 0;
    BESST_CUDA_TRY(ctx, ctx->tile_aggs.ensure(sizeof(Agg) * (size_t)(n_tiles + n_chunks + 2)));
    BESST_CUDA_TRY(ctx, ctx->tile_state.ensure(sizeof(u64) * (size_t)(n_tiles + 2) + sizeof(ChunkIn) * (size_t)(n_chunks + 1)));
    BESST_CUDA_TRY(ctx, ctx->scratch_tuples.ensure(sizeof(besst_link_tuple) * (size_t)(n_tiles > 0 ? n_tiles : 1) * WT));
    BESST_CUDA_TRY(ctx, ctx->block_tile0.ensure(sizeof(u32) * (size_t)(n_tiles / 16 + 4)));   // one entry per 2048 accepted tuples
    if (ctx->fishy_cap == 0) ctx->fishy_cap = n / 8 + 4096;
    BESST_CUDA_TRY(ctx, ctx->fishy_keys.ensure(sizeof(u64) * (size_t)ctx->fishy_cap));
    BESST_CUDA_TRY(ctx, cudaMemsetAsync(ctx->aligned.p, 0, sizeof(u64) * (c1 + BESST_N_COUNTERS + 8), ctx->stream));
    return BESST_OK;
}

// K1 over records [r0, r1) of `rec` (device pointers to the WHOLE batch); r0 must be a multiple of 128
int besst_extract_slice(besst_ctx* ctx, const besst_lib_params& p, const DeviceRecords& rec, int64_t r0, int64_t r1) {
    if (r1 <= r0) return BESST_OK;
    if (r0 % WT) { ctx->err = "extract slice: unaligned start"; return BESST_E_INVALID; }
    const int64_t tile0 = r0 / WT, n_tiles = tiles_of(r1 - r0);
    u64* counters = ctx->counters.as<u64>();
    u64* globals = counters + BESST_N_COUNTERS;
    K1Params P;
    P.rec.n = r1 - r0;
    P.rec.tid = rec.tid + r0; P.rec.mtid = rec.mtid + r0; P.rec.pos = rec.pos + r0; P.rec.mpos = rec.mpos + r0;
    const bool packed = rec.packed != nullptr;
    P.rec.qlen = packed ? nullptr : rec.qlen + r0; P.rec.flag = packed ? nullptr : rec.flag + r0; P.rec.mapq = packed ? nullptr : rec.mapq + r0;
    P.rec.packed = packed ? rec.packed + r0 : nullptr; P.rec.tlen = nullptr;
    bool vec = true;
    const void* ptrs[] = {P.rec.tid, P.rec.mtid, P.rec.pos, P.rec.mpos, P.rec.qlen, P.rec.flag, P.rec.mapq, P.rec.packed};
    for (const void* q : ptrs) vec = vec && ((reinterpret_cast<uintptr_t>(q) & 15u) == 0);
    const bool int_rl = p.read_len >= 0 && p.read_len < 1e9 && p.read_len == (double)(long long)p.read_len;
    P.rows = ctx->rows_packed.as<int4>();
    P.n_contigs = (int)ctx->n_contigs;
    P.orientation = p.orientation; P.min_mapq = p.min_mapq; P.detect_dup = p.detect_duplicate;
    P.extend = p.extend_paths; P.scoring = !p.no_score;
    P.read_len = p.read_len;
    P.read_len_i = int_rl ? (int)p.read_len : 0;
    {   // integer form of the acceptance threshold (exact for integer sums; NaN never accepts)
        const double t = p.ins_size_threshold;
        if (t != t) P.threshold_i = -(1ll << 62);
        else if (t >= 4e18) P.threshold_i = (1ll << 62);
        else if (t <= -4e18) P.threshold_i = -(1ll << 62);
        else P.threshold_i = (long long)ceil(t);
    }
    P.scratch = ctx->scratch_tuples.as<besst_link_tuple>() + tile0 * WT;
    P.aggs = ctx->tile_aggs.as<Agg>() + tile0;
    P.fishy = ctx->fishy_keys.as<u64>(); P.fishy_cap = ctx->fishy_cap;
    P.aligned = ctx->aligned.as<u64>();
    P.counters = counters;
    P.globals = globals;
    P.n_tiles = n_tiles;

    typedef void (*K1Fn)(const K1Params, const int);
    const K1Fn k1 = packed ? (int_rl ? k_extract_links_tma<true, true> : k_extract_links_tma<false, true>)
                           : (int_rl ? k_extract_links_tma<true, false> : k_extract_links_tma<false, false>);
    const size_t smem = sizeof(WarpSmem) * K1T_WARPS;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1, K1T_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)ctx->sm_count * per_sm;
    const long long max_grid = (n_tiles + (long long)K1T_WARPS * K1T_BATCH - 1) / ((long long)K1T_WARPS * K1T_BATCH);
    if (grid > max_grid) grid = max_grid > 0 ? max_grid : 1;
    BESST_CUDA_TRY(ctx, cudaMemsetAsync(globals + 2, 0, sizeof(u64), ctx->stream));   // the tile ticket
    {
        KTimer kt(ctx, BESST_K_EXTRACT);
        k1<<<(unsigned)grid, K1T_THREADS, smem, ctx->stream>>>(P, vec ? 1 : 0);
    }
    BESST_CUDA_TRY(ctx, cudaGetLastError());
    return BESST_OK;
}

// scan of the tile aggregates, sizes to the host, compaction.  *overflow: the fishy-key buffer was
// too small (it has been grown; run begin/slice/finish again)
int besst_extract_finish(besst_ctx* ctx, const besst_lib_params& p, int64_t n, bool* overflow) {
    *overflow = false;
    const int64_t n_tiles = tiles_of(n);
    const int64_t n_chunks = (n_tiles + SC_THREADS - 1) / SC_THREADS;
    Agg* aggs = ctx->tile_aggs.as<Agg>();
    Agg* chunk_aggs = aggs + n_tiles + 1;
    u64* tile_off = ctx->tile_state.as<u64>();
    ChunkIn* chunk_in = reinterpret_cast<ChunkIn*>(tile_off + n_tiles + 2);
    u64* counters = ctx->counters.as<u64>();
    u64* globals = counters + BESST_N_COUNTERS;
    if (n_tiles > 0) {
        { KTimer kt(ctx, BESST_K_TILE_SCAN); k_tile_reduce<<<(unsigned)n_chunks, SC_THREADS, 0, ctx->stream>>>(aggs, n_tiles, chunk_aggs, p.detect_duplicate); }
        { KTimer kt(ctx, BESST_K_TILE_SCAN); k_chunk_resolve<<<1, SC_THREADS, 0, ctx->stream>>>(chunk_aggs, n_chunks, chunk_in, tile_off + n_tiles, p.halo_prev_obs1, p.halo_prev_obs2, p.detect_duplicate, counters, globals); }
        { KTimer kt(ctx, BESST_K_TILE_SCAN); k_tile_offsets<<<(unsigned)n_chunks, SC_THREADS, 0, ctx->stream>>>(aggs, n_tiles, chunk_in, tile_off, p.detect_duplicate, ctx->block_tile0.as<u32>()); }
        BESST_CUDA_TRY(ctx, cudaGetLastError());
    } else {
        const int64_t halo[2] = {p.halo_prev_obs1, p.halo_prev_obs2};
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(counters + BESST_CNT_LAST_OBS1, halo, sizeof(halo), cudaMemcpyHostToDevice, ctx->stream));
    }
    u64* const g = reinterpret_cast<u64*>(ctx->host_scalars());   // pinned: no bounce buffer on the size read-backs
    if (!g) { ctx->err = "pinned host scratch allocation failed"; return BESST_E_NOMEM; }
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(g, globals, 16, cudaMemcpyDeviceToHost, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_tuples = (int64_t)g[0];
    ctx->n_fishy_keys = (int64_t)g[1];
    if (ctx->n_fishy_keys > ctx->fishy_cap) {
        ctx->fishy_cap = ctx->n_fishy_keys + 4096;
        *overflow = true;
        return BESST_OK;
    }
    // the accepted tuples stay in their tile-local scratch runs: the run-merge bucket reads them from
    // there (tile_off says where); the BAM-ordered array is only materialised on demand (besst_ensure_tuples)
    ctx->n_rec_tiles = n_tiles;
    ctx->tuples_valid = false;
    ctx->have_links = true;
    return BESST_OK;
}

// materialise the BAM-ordered tuple array (tuple-level multi-GPU exchange, radix fallback, ABI accessors)
int besst_ensure_tuples(besst_ctx* ctx) {
    if (!ctx->have_links) { ctx->err = "no extracted links"; return BESST_E_STATE; }
    if (ctx->tuples_valid) return BESST_OK;
    BESST_CUDA_TRY(ctx, ctx->tuples.ensure(sizeof(besst_link_tuple) * (size_t)(ctx->n_tuples > 0 ? ctx->n_tuples : 1)));
    if (ctx->n_tuples > 0) {
        const long long n_tiles = ctx->n_rec_tiles;
        long long cgrid = (long long)ctx->sm_count * 8;
        const long long cmax = (n_tiles + 7) / 8;
        if (cgrid > cmax) cgrid = cmax;
        KTimer kt(ctx, BESST_K_COMPACT);
        k_compact_tuples<<<(unsigned)cgrid, 256, 0, ctx->stream>>>(ctx->scratch_tuples.as<besst_link_tuple>(), ctx->tile_state.as<u64>(), n_tiles,
                                                                   ctx->tuples.as<besst_link_tuple>());
        BESST_CUDA_TRY(ctx, cudaGetLastError());
    }
    ctx->tuples_valid = true;
    return BESST_OK;
}

int besst_launch_extract(besst_ctx* ctx, const besst_lib_params& p, const DeviceRecords& rec) {
    for (int attempt = 0; attempt < 3; ++attempt) {
        int rc = besst_extract_begin(ctx, p, rec.n);
        if (rc) return rc;
        rc = besst_extract_slice(ctx, p, rec, 0, rec.n);
        if (rc) return rc;
        bool overflow = false;
        rc = besst_extract_finish(ctx, p, rec.n, &overflow);
        if (rc) return rc;
        if (!overflow) return BESST_OK;
    }
    ctx->err = "link extraction: fishy-key capacity retry failed";
    return BESST_E_STATE;
}
