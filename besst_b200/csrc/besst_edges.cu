// K4-K6  sorted link tuples -> CSR edge list with link statistics, KS span
// score and ML gap estimate.  Compiled with -fmad=false so that the fp64
// arithmetic rounds exactly like the reference's Python (and the C oracle).
//
//  K4  segment heads -> row_ptr; one warp per edge reduces its segment:
//      nr_links, obs = sum(o1+o2), obs_sq = sum((o1+o2)^2), first-appearance
//      index, per-scaffold observation lists in BAM order
//      (CreateEdge, CreateGraph.py:842-862), fishy count (:141-163).
//  K5  two-sample KS statistic = scipy.stats.ks_2samp(...).statistic as used at
//      CreateGraph.py:582-606: the two observation lists of every large-large
//      edge are sorted by two device-wide radix sorts of (edge, value) keys and
//      compared by a co-ranking walk, load-balanced over links.
//  K6  GapEstimator bisection + tr_sk_std_dev (mathstats param_est, call sites
//      CreateGraph.py:537,555) with the four erf/exp arguments of g(d) spread
//      over the 4 lanes of a quad and combined with __shfl_sync; then the score
//      (CreateGraph.py:603-614).  No tensor cores: there is no contraction here.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "besst_internal.cuh"

int besst_radix_sort_tuples(besst_ctx* ctx, const besst_link_tuple* tuples, int bv, uint64_t* keys_a, uint64_t* keys_b,
                            uint32_t* val_a, uint32_t* val_b, int64_t n, int* result_in_b);
int besst_radix_sort_tuples_packed(besst_ctx* ctx, const besst_link_tuple* tuples, int bv, int idx_bits, uint64_t* keys_a,
                                   uint64_t* keys_b, int64_t n, int* result_in_b);
int besst_radix_sort_keys32(besst_ctx* ctx, uint32_t* keys_a, uint32_t* keys_b, int64_t n, int key_bits, int* result_in_b);

namespace {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int HB_THREADS = 256;
constexpr int HB_ITEMS = 8;
constexpr int HB_TILE = HB_THREADS * HB_ITEMS;

// ---- segment heads ---------------------------------------------------------------
// sorted words are (edge key << idx_bits) | index when the index travels packed in the key (idx_bits > 0)
__global__ void __launch_bounds__(HB_THREADS) k_head_count(const u64* __restrict__ keys, long long n, int idx_bits, u32* block_sums) {
    __shared__ u32 s_w[HB_THREADS / 32];
    const long long base = (long long)blockIdx.x * HB_TILE;
    u32 c = 0;
#pragma unroll
    for (int i = 0; i < HB_ITEMS; ++i) {
        const long long j = base + i * HB_THREADS + threadIdx.x;
        if (j < n) c += (j == 0 || (keys[j] >> idx_bits) != (keys[j - 1] >> idx_bits)) ? 1u : 0u;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0;
        for (int w = 0; w < HB_THREADS / 32; ++w) t += s_w[w];
        block_sums[blockIdx.x] = t;
    }
}

// single-CTA exclusive scan of the block sums; total -> block_sums[n_blocks]
__global__ void __launch_bounds__(1024) k_scan_blocks(u32* block_sums, int n_blocks) {
    __shared__ u32 s_w[32];
    __shared__ u32 s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const u32 v = i < n_blocks ? block_sums[i] : 0;
        u32 incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            u32 w = s_w[lane];
            u32 wi = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const u32 t = __shfl_up_sync(0xffffffffu, wi, off);
                if (lane >= off) wi += t;
            }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const u32 carry = s_carry;
        if (i < n_blocks) block_sums[i] = carry + s_w[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_w[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[n_blocks] = s_carry;
}

__global__ void __launch_bounds__(HB_THREADS)
    k_head_write(const u64* __restrict__ keys, long long n, int idx_bits, const u32* __restrict__ block_sums, long long* row_ptr) {
    __shared__ u32 s_w[HB_THREADS / 32];
    const long long base = (long long)blockIdx.x * HB_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // blocked arrangement so that heads keep their order
    bool head[HB_ITEMS];
    u32 c = 0;
#pragma unroll
    for (int i = 0; i < HB_ITEMS; ++i) {
        const long long j = base + (long long)threadIdx.x * HB_ITEMS + i;
        head[i] = j < n && (j == 0 || (keys[j] >> idx_bits) != (keys[j - 1] >> idx_bits));
        c += head[i] ? 1u : 0u;
    }
    u32 incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    u32 wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += s_w[w];
    u32 pos = block_sums[blockIdx.x] + wbase + incl - c;
#pragma unroll
    for (int i = 0; i < HB_ITEMS; ++i)
        if (head[i]) row_ptr[pos++] = base + (long long)threadIdx.x * HB_ITEMS + i;
    if (blockIdx.x == 0 && threadIdx.x == 0) row_ptr[block_sums[gridDim.x]] = n;
}

// ---- fishy keys: (u<<32)|v -> (u<<bv)|v so that they sort in 2*bv bits ---------------
__global__ void k_fishy_rekey(const u64* in, u64* out, long long n, int bv) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const u64 k = in[i];
        out[i] = ((k >> 32) << bv) | (k & 0xffffffffull);
    }
}

__device__ __forceinline__ long long lower_bound_u64(const u64* a, long long n, u64 key) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- K4: one warp per edge ----------------------------------------------------------
struct EdgeArrays {
    u32 *u, *v;
    int* nr;
    long long *obs, *obs_sq, *first, *row_ptr;
    int* gap;
    double *score, *ks, *sd_obs, *sd_model;
    int* fishy;
    unsigned char* flags;
    int *obs_u, *obs_v;
    long long* sum_u;   // internal: sum of obs_u per edge
    int* max_v;         // internal: max of obs_v per edge
};

__global__ void __launch_bounds__(256)
    k_edge_reduce(EdgeArrays E, long long n_edges, const besst_link_tuple* __restrict__ tuples,
                  const u32* __restrict__ sorted_idx, const u64* __restrict__ sorted_keys, int bv, int idx_bits,
                  const u64* __restrict__ fishy_sorted, long long n_fishy, u32 n_large2, int scoring) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long e = warp_global; e < n_edges; e += n_warps) {
        const long long b = E.row_ptr[e], t = E.row_ptr[e + 1];
        long long s = 0, sq = 0, su = 0;
        int mv = -2147483647 - 1;
        for (long long j = b + lane; j < t; j += 32) {
            const u32 idx = idx_bits ? (u32)(__ldg(sorted_keys + j) & ((1ull << idx_bits) - 1ull)) : __ldg(sorted_idx + j);
            const int4 tp = __ldg(reinterpret_cast<const int4*>(tuples + idx));
            E.obs_u[j] = tp.z;
            E.obs_v[j] = tp.w;
            const long long o = (long long)tp.z + (long long)tp.w;
            s += o;
            sq += o * o;
            su += tp.z;
            mv = tp.w > mv ? tp.w : mv;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, off);
            sq += __shfl_xor_sync(0xffffffffu, sq, off);
            su += __shfl_xor_sync(0xffffffffu, su, off);
            const int o = __shfl_xor_sync(0xffffffffu, mv, off);
            mv = o > mv ? o : mv;
        }
        if (lane == 0) {
            const u64 word = sorted_keys[b];
            const u64 key = word >> idx_bits;
            const u32 u = (u32)(key >> bv), v = (u32)(key & ((1ull << bv) - 1ull));
            E.u[e] = u;
            E.v[e] = v;
            E.nr[e] = (int)(t - b);
            E.obs[e] = s;
            E.obs_sq[e] = sq;
            E.first[e] = idx_bits ? (long long)(word & ((1ull << idx_bits) - 1ull)) : (long long)sorted_idx[b];
            long long f = 0;
            if (n_fishy > 0) {
                const long long lo = lower_bound_u64(fishy_sorted, n_fishy, key);
                const long long hi = lower_bound_u64(fishy_sorted, n_fishy, key + 1);
                f = hi - lo;
            }
            E.fishy[e] = (int)f;
            const bool ll = u < n_large2 && v < n_large2;
            E.flags[e] = ll ? BESST_EDGE_LL : 0;
            E.gap[e] = 0;
            E.sum_u[e] = su;
            E.max_v[e] = mv;
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            E.score[e] = nan; E.sd_obs[e] = nan; E.sd_model[e] = nan;
            E.ks[e] = (ll && scoring) ? 0.0 : nan;   // k_ks_eval accumulates the maximum into it
        }
    }
}

// ---- K3': run-merge bucket ---------------------------------------------------------------------------
// The tuple stream is in BAM order, i.e. sorted by the record's contig: the links of one edge sit in
// a few dense stretches of the stream (the stretch of either contig).  A device-wide radix sort
// ignores that and moves every link five times.  Instead:
//   k_group_blocks  every CTA takes 2048 consecutive tuples, gives each distinct (u,v) a dense local
//                   id through a shared-memory hash table, ranks the tuples stably by id (per-warp
//                   __match_any_sync multisplit + prefix over warps) and writes the block's
//                   observations grouped by edge, BAM order kept, plus one RUN descriptor per
//                   (block, edge): key, where the run starts, its length, its first BAM ordinal;
//   sort            the run descriptors (a few per contig, ~0.6 % of the links at config 3) are radix
//                   sorted by (u, v, block): runs of one edge become adjacent, in BAM order;
//   k_run_*         heads / prefix sums over the sorted runs -> edges, row_ptr, where each run goes;
//   k_edge_gather   one warp per edge copies its runs to their CSR position with coalesced 8-byte
//                   accesses and reduces nr_links / obs / obs_sq (CreateEdge, CreateGraph.py:842-862).
// Each link is read twice and written twice.  Input that is not locally ordered makes many runs: the
// caller falls back to the radix bucket when a block holds more than GB_DMAX edges or the runs
// exceed a fraction of the links.
constexpr int GB_THREADS = 256;
constexpr int GB_WARPS = GB_THREADS / 32;
constexpr int GB_ITEMS = 8;
constexpr int GB_TILE = GB_THREADS * GB_ITEMS;   // 2048 tuples per block
constexpr int GB_DMAX = 512;                     // distinct edges per block
constexpr int GB_HT = 2 * GB_DMAX;               // hash slots (load <= 0.5 for the blocks this kernel accepts)
constexpr int GB_PROBES = 96;                    // longer probe sequences mean more than GB_DMAX edges (or bad luck): fall back
constexpr u64 GB_EMPTY = ~0ull;

struct GroupSmem {
    union {
        u64 ht[GB_HT];          // hash table of (u << 32 | v) ...
        int2 stage[GB_TILE];    // ... later the observations in grouped order
    };
    static_assert(sizeof(u64) * GB_HT <= sizeof(int2) * GB_TILE, "union sizes");
    unsigned short id_of_slot[GB_HT];
    unsigned short warp_hist[GB_WARPS][GB_DMAX];
    u64 key_of_id[GB_DMAX];
    unsigned short start[GB_DMAX + 2];
    u32 scan_w[GB_WARPS];
    u32 n_ids, run_base, overflow;
    // FROM_SCRATCH: the non-empty record tiles overlapping this block
    u32 tile_id[GB_TILE + 2];      // tile number | drop-first flag << 31
    u32 tile_start[GB_TILE + 2];   // ordinal of the tile's first kept tuple, relative to the block
    u32 n_list, scan_tile0;
};
constexpr u64 GB_OFF_DROP = 1ull << 63;   // tile_off flag (besst_links.cu): skip the tile's first scratch tuple

__device__ __forceinline__ u32 gb_insert(u64* ht, u64 key, u32* overflow) {
    u32 h = ((u32)(key >> 32) * 0x9E3779B1u) ^ ((u32)key * 0x85EBCA77u);
    h = (h >> 12) & (GB_HT - 1);
    for (int probe = 0; probe < GB_PROBES; ++probe) {
        const u64 cur = ht[h];
        if (cur == key) return h;
        if (cur == GB_EMPTY) {
            const u64 old = atomicCAS(reinterpret_cast<unsigned long long*>(ht + h), GB_EMPTY, key);
            if (old == GB_EMPTY || old == key) return h;
        }
        h = (h + 1) & (GB_HT - 1);
    }
    *overflow = 1;
    return 0;
}

// gstate: [0] = runs so far, [1] = overflow flags (1: a block with too many edges, 2: run capacity)
// FROM_SCRATCH: `tuples` is K1's scratch (128 slots per record tile, the kept tuples of tile t at
// t * 128 + drop .. ) and tile_off[t] the ordinal of tile t's first kept tuple: the block finds the tiles
// that overlap its 2048 ordinals and reads them in place -- no compacted copy of the tuple stream.
#ifndef BESST_GB_MIN_CTAS
#define BESST_GB_MIN_CTAS 4
#endif
template <bool FROM_SCRATCH>
__global__ void __launch_bounds__(GB_THREADS, BESST_GB_MIN_CTAS)
    k_group_blocks(const besst_link_tuple* __restrict__ tuples, long long n, int bv, int block_bits, int2* __restrict__ grouped,
                   u64* __restrict__ run_key, u32* __restrict__ run_val, u32* __restrict__ run_start, u32* __restrict__ run_cnt,
                   u32* __restrict__ run_first, u32* gstate, u32 run_cap, const u64* __restrict__ tile_off, long long n_tiles,
                   const u32* __restrict__ block_tile0) {
    extern __shared__ __align__(16) unsigned char gb_smem[];
    GroupSmem& S = *reinterpret_cast<GroupSmem*>(gb_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    const long long base = (long long)blockIdx.x * GB_TILE;
    const int count = (n - base < GB_TILE) ? (int)(n - base) : GB_TILE;

    for (int i = threadIdx.x; i < GB_HT / 2; i += GB_THREADS) reinterpret_cast<ulonglong2*>(S.ht)[i] = make_ulonglong2(GB_EMPTY, GB_EMPTY);
    for (int i = threadIdx.x; i < GB_WARPS * GB_DMAX / 8; i += GB_THREADS) reinterpret_cast<uint4*>(&S.warp_hist[0][0])[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { S.n_ids = 0; S.overflow = 0; S.n_list = 0; }
    if (FROM_SCRATCH && threadIdx.x == 0) S.scan_tile0 = __ldg(block_tile0 + blockIdx.x);   // the tile holding ordinal `base` (k_tile_offsets)
    __syncthreads();
    if (FROM_SCRATCH) {   // list of the non-empty tiles overlapping [base, base + count)
        const long long end = base + count;
        for (long long t0 = S.scan_tile0;; t0 += GB_THREADS) {
            const long long t = t0 + threadIdx.x;
            long long a = end, b = end;
            u64 raw = 0;
            if (t < n_tiles) {
                raw = __ldg(tile_off + t);
                a = (long long)(raw & ~GB_OFF_DROP);
                b = (long long)(__ldg(tile_off + t + 1) & ~GB_OFF_DROP);
            }
            const bool keep = b > a && a < end && b > base;
            const u32 bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) S.scan_w[warp] = (u32)__popc(bal);
            __syncthreads();
            u32 pos = S.n_list;
            for (int w = 0; w < warp; ++w) pos += S.scan_w[w];
            if (keep) {
                const u32 k = pos + (u32)__popc(bal & lt_mask);
                S.tile_id[k] = (u32)t | ((raw & GB_OFF_DROP) ? 0x80000000u : 0u);
                S.tile_start[k] = (u32)(a - base + 0x40000000ll);   // biased: the first tile may start before the block
            }
            const bool more = __syncthreads_or(t == t0 + GB_THREADS - 1 && a < end && t + 1 < n_tiles);
            if (threadIdx.x == 0) {
                u32 tot = 0;
                for (int w = 0; w < GB_WARPS; ++w) tot += S.scan_w[w];
                S.n_list += tot;
            }
            __syncthreads();
            if (!more) break;
        }
    }

    // ---- phase 1: load, warp-level dedup of the keys, hash insert by one lane per distinct key ----
    int2 obs[GB_ITEMS];
    int list_pos = 0;
    u32 meta[GB_ITEMS];   // slot (12) | lanes of the same key before me (5) << 12 | same-key lanes - 1 (5) << 17 | leader (5) << 22
#pragma unroll
    for (int i = 0; i < GB_ITEMS; ++i) {
        const int p = warp * (32 * GB_ITEMS) + i * 32 + lane;
        const bool valid = p < count;
        const u32 vmask = __ballot_sync(0xffffffffu, valid);
        meta[i] = 0;
        obs[i] = make_int2(0, 0);
        if (valid) {
            long long src = base + p;
            if (FROM_SCRATCH) {   // last listed tile starting at or before ordinal p
                const u32 want = (u32)p + 0x40000000u;
                int lo = list_pos;
                if (i == 0) {   // binary search once; the following items continue from the previous tile
                    int hi = (int)S.n_list;   // tile_start[lo] <= want < tile_start[hi]
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (S.tile_start[mid] <= want) lo = mid; else hi = mid;
                    }
                } else {
                    const int last = (int)S.n_list - 1;
                    while (lo < last && S.tile_start[lo + 1] <= want) ++lo;
                }
                list_pos = lo;
                const u32 id = S.tile_id[lo];
                src = (long long)(id & 0x7fffffffu) * 128 + (long long)(want - S.tile_start[lo]) + (id >> 31);
            }
            const int4 t = __ldg(reinterpret_cast<const int4*>(tuples + src));
            obs[i] = make_int2(t.z, t.w);
            const u64 key = ((u64)(u32)t.x << 32) | (u32)t.y;
            const u32 peers = __match_any_sync(vmask, key);
            const int leader = __ffs(peers) - 1;
            u32 slot = 0;
            if (lane == leader) slot = gb_insert(S.ht, key, &S.overflow);
            slot = __shfl_sync(vmask, slot, leader);
            meta[i] = slot | ((u32)__popc(peers & lt_mask) << 12) | ((u32)(__popc(peers) - 1) << 17) | ((u32)leader << 22);
        }
    }
    __syncthreads();

    // ---- phase 2: dense local ids in slot order --------------------------------------------------
    {
        constexpr int PER = GB_HT / GB_THREADS;   // slots k * GB_THREADS + thread: conflict-free
        u32 used = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) used |= (S.ht[k * GB_THREADS + threadIdx.x] != GB_EMPTY ? 1u : 0u) << k;
        const u32 c = (u32)__popc(used);
        u32 incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) S.scan_w[warp] = incl;
        __syncthreads();
        u32 id = incl - c;
        u32 total = 0;
#pragma unroll
        for (int w = 0; w < GB_WARPS; ++w) {
            if (w < warp) id += S.scan_w[w];
            total += S.scan_w[w];
        }
#pragma unroll
        for (int k = 0; k < PER; ++k)
            if (used >> k & 1u) {
                const int s = k * GB_THREADS + threadIdx.x;
                S.id_of_slot[s] = (unsigned short)id;
                if (id < GB_DMAX) S.key_of_id[id] = S.ht[s];
                ++id;
            }
        if (threadIdx.x == 0) {
            S.n_ids = total;
            if (total > GB_DMAX || S.overflow) { S.overflow = 1; atomicOr(gstate + 1, 1u); }
            else S.run_base = atomicAdd(gstate, total);
        }
    }
    __syncthreads();
    if (S.overflow) return;   // the caller falls back to the radix bucket
    const int D = (int)S.n_ids;

    // ---- phase 3: stable rank inside the warp, items in BAM order ------------------------------------
#pragma unroll
    for (int i = 0; i < GB_ITEMS; ++i) {
        const int p = warp * (32 * GB_ITEMS) + i * 32 + lane;
        const bool valid = p < count;
        const u32 vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const u32 m = meta[i];
            const u32 id = S.id_of_slot[m & 0xfffu];
            const int leader = (int)(m >> 22);
            u32 pre = 0;
            if (lane == leader) {
                pre = S.warp_hist[warp][id];
                S.warp_hist[warp][id] = (unsigned short)(pre + ((m >> 17) & 31u) + 1u);
            }
            pre = __shfl_sync(vmask, pre, leader);
            meta[i] = id | ((pre + ((m >> 12) & 31u)) << 9);   // id (9) | rank within the warp (<= 255)
        }
        __syncwarp();
    }
    __syncthreads();

    // ---- phase 4: prefix over warps, starts of the ids, run descriptors ----------------------------------
    {
        const int id0 = 2 * threadIdx.x, id1 = id0 + 1;   // GB_DMAX == 2 * GB_THREADS
        u32 tot0 = 0, tot1 = 0;
        if (id0 < D) {
#pragma unroll
            for (int w = 0; w < GB_WARPS; ++w) { const u32 t = S.warp_hist[w][id0]; S.warp_hist[w][id0] = (unsigned short)tot0; tot0 += t; }
        }
        if (id1 < D) {
#pragma unroll
            for (int w = 0; w < GB_WARPS; ++w) { const u32 t = S.warp_hist[w][id1]; S.warp_hist[w][id1] = (unsigned short)tot1; tot1 += t; }
        }
        const u32 c = tot0 + tot1;
        u32 incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) S.scan_w[warp] = incl;
        __syncthreads();
        u32 st = incl - c;
#pragma unroll
        for (int w = 0; w < GB_WARPS; ++w)
            if (w < warp) st += S.scan_w[w];
        const u32 rb = S.run_base;
        if (id0 < D) {
            S.start[id0] = (unsigned short)st;
            const u32 r = rb + (u32)id0;
            if (r < run_cap) {
                const u64 k = S.key_of_id[id0];
                run_key[r] = (((k >> 32) << bv | (k & 0xffffffffull)) << block_bits) | (u64)blockIdx.x;
                run_val[r] = r; run_start[r] = (u32)(base + st); run_cnt[r] = tot0;
            } else atomicOr(gstate + 1, 2u);
        }
        if (id1 < D) {
            S.start[id1] = (unsigned short)(st + tot0);
            const u32 r = rb + (u32)id1;
            if (r < run_cap) {
                const u64 k = S.key_of_id[id1];
                run_key[r] = (((k >> 32) << bv | (k & 0xffffffffull)) << block_bits) | (u64)blockIdx.x;
                run_val[r] = r; run_start[r] = (u32)(base + st + tot0); run_cnt[r] = tot1;
            } else atomicOr(gstate + 1, 2u);
        }
    }
    __syncthreads();

    // ---- phase 5: final position inside the block; the observations go through shared memory ----------
    {
        const u32 rb = S.run_base;
#pragma unroll
        for (int i = 0; i < GB_ITEMS; ++i) {
            const int p = warp * (32 * GB_ITEMS) + i * 32 + lane;
            if (p < count) {
                const u32 id = meta[i] & 511u;
                const u32 in_run = (u32)S.warp_hist[warp][id] + (meta[i] >> 9);
                S.stage[S.start[id] + in_run] = obs[i];
                if (in_run == 0 && rb + id < run_cap) run_first[rb + id] = (u32)(base + p);   // the run's first link, BAM order
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < count; j += GB_THREADS) grouped[base + j] = S.stage[j];
}

// ---- heads and prefix sums over the sorted runs: packed (heads << 32 | links) ----------------------------
constexpr int RS_T = 256, RS_I = 8, RS_TILE = RS_T * RS_I;

__global__ void __launch_bounds__(RS_T)
    k_run_count(const u64* __restrict__ keys, const u32* __restrict__ vals, const u32* __restrict__ run_cnt, long long R,
                int block_bits, u64* block_sums) {
    __shared__ u64 s_w[RS_T / 32];
    const long long base = (long long)blockIdx.x * RS_TILE;
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < RS_I; ++i) {
        const long long r = base + i * RS_T + threadIdx.x;
        if (r < R) {
            const bool head = r == 0 || (keys[r] >> block_bits) != (keys[r - 1] >> block_bits);
            c += ((u64)(head ? 1u : 0u) << 32) | (u64)__ldg(run_cnt + __ldg(vals + r));
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (int w = 0; w < RS_T / 32; ++w) t += s_w[w];
        block_sums[blockIdx.x] = t;
    }
}

// single-CTA exclusive scan of packed sums; total -> block_sums[n_blocks]
__global__ void __launch_bounds__(1024) k_scan_blocks64(u64* block_sums, int n_blocks) {
    __shared__ u64 s_w[32];
    __shared__ u64 s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const u64 v = i < n_blocks ? block_sums[i] : 0;
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
        if (i < n_blocks) block_sums[i] = carry + pre + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[n_blocks] = s_carry;
}

// per sorted run: where its links go (run_off), where they come from (run_src), how many (run_len);
// per edge (at its first run): row_ptr, first run, u, v, first BAM ordinal
__global__ void __launch_bounds__(RS_T)
    k_run_write(const u64* __restrict__ keys, const u32* __restrict__ vals, const u32* __restrict__ run_cnt,
                const u32* __restrict__ run_start, const u32* __restrict__ run_first, long long R, int block_bits, int bv,
                const u64* __restrict__ block_sums, u32* run_off, u32* run_src, u32* run_len, long long* row_ptr,
                u32* edge_run_ptr, u32* edge_u, u32* edge_v, long long* edge_first, long long n_links) {
    __shared__ u64 s_w[RS_T / 32];
    const long long base = (long long)blockIdx.x * RS_TILE + (long long)threadIdx.x * RS_I;   // blocked: order kept
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 item[RS_I];
    u32 val[RS_I];
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < RS_I; ++i) {
        const long long r = base + i;
        item[i] = 0; val[i] = 0;
        if (r < R) {
            const bool head = r == 0 || (keys[r] >> block_bits) != (keys[r - 1] >> block_bits);
            val[i] = __ldg(vals + r);
            item[i] = ((u64)(head ? 1u : 0u) << 32) | (u64)__ldg(run_cnt + val[i]);
        }
        c += item[i];
    }
    u64 incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    u64 pos = block_sums[blockIdx.x] + incl - c;
    for (int w = 0; w < warp; ++w) pos += s_w[w];
#pragma unroll
    for (int i = 0; i < RS_I; ++i) {
        const long long r = base + i;
        if (r < R) {
            const u32 link_off = (u32)pos, e = (u32)(pos >> 32);   // exclusive: links before this run, heads before it
            run_off[r] = link_off;
            run_src[r] = __ldg(run_start + val[i]);
            run_len[r] = (u32)item[i];
            if (item[i] >> 32) {
                const u64 key = keys[r] >> block_bits;
                row_ptr[e] = (long long)link_off;
                edge_run_ptr[e] = (u32)r;
                edge_u[e] = (u32)(key >> bv);
                edge_v[e] = (u32)(key & ((1ull << bv) - 1ull));
                edge_first[e] = (long long)__ldg(run_first + val[i]);
            }
        }
        pos += item[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const u32 E = (u32)(block_sums[gridDim.x] >> 32);
        row_ptr[E] = n_links;
        edge_run_ptr[E] = (u32)R;
    }
}

// K4': one warp per edge: copy its runs to CSR order and reduce
// PACK16: the observations arrive as one 32-bit word per link, obs_u | obs_v << 16 (multi-GPU exchange when
// ins_size_threshold <= 65535: every accepted observation is below the threshold, CreateGraph.py:840)
template <bool PACK16>
__global__ void __launch_bounds__(256)
    k_edge_gather(EdgeArrays E, long long n_edges, const int2* __restrict__ grouped, const u32* __restrict__ edge_run_ptr,
                  const u32* __restrict__ run_off, const u32* __restrict__ run_src, const u32* __restrict__ run_len, int bv,
                  const u64* __restrict__ fishy_sorted, long long n_fishy, u32 n_large2, int scoring) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long e = warp_global; e < n_edges; e += n_warps) {
        const u32 r0 = __ldg(edge_run_ptr + e), r1 = __ldg(edge_run_ptr + e + 1);
        long long s = 0, sq = 0, su = 0;
        int mv = -2147483647 - 1;
        for (u32 r = r0; r < r1; ++r) {
            const u32 src = __ldg(run_src + r), dst = __ldg(run_off + r), len = __ldg(run_len + r);
            for (u32 k = lane; k < len; k += 32) {
                int2 o;
                if (PACK16) {
                    const u32 w = __ldg(reinterpret_cast<const u32*>(grouped) + src + k);
                    o = make_int2((int)(w & 0xffffu), (int)(w >> 16));
                } else {
                    o = __ldg(grouped + src + k);
                }
                E.obs_u[dst + k] = o.x;
                E.obs_v[dst + k] = o.y;
                const long long t = (long long)o.x + (long long)o.y;
                s += t;
                sq += t * t;
                su += o.x;
                mv = o.y > mv ? o.y : mv;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, off);
            sq += __shfl_xor_sync(0xffffffffu, sq, off);
            su += __shfl_xor_sync(0xffffffffu, su, off);
            const int o = __shfl_xor_sync(0xffffffffu, mv, off);
            mv = o > mv ? o : mv;
        }
        if (lane == 0) {
            const u32 u = E.u[e], v = E.v[e];
            const u64 key = ((u64)u << bv) | v;
            E.nr[e] = (int)(E.row_ptr[e + 1] - E.row_ptr[e]);
            E.obs[e] = s;
            E.obs_sq[e] = sq;
            long long f = 0;
            if (n_fishy > 0) {
                const long long lo = lower_bound_u64(fishy_sorted, n_fishy, key);
                const long long hi = lower_bound_u64(fishy_sorted, n_fishy, key + 1);
                f = hi - lo;
            }
            E.fishy[e] = (int)f;
            const bool ll = u < n_large2 && v < n_large2;
            E.flags[e] = ll ? BESST_EDGE_LL : 0;
            E.gap[e] = 0;
            E.sum_u[e] = su;
            E.max_v[e] = mv;
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            E.score[e] = nan; E.sd_obs[e] = nan; E.sd_model[e] = nan;
            E.ks[e] = (ll && scoring) ? 0.0 : nan;   // k_ks_eval accumulates the maximum into it
        }
    }
}

// ---- mathstats restatement on the device (see oracle/besst_oracle.c) ------------------
__device__ __forceinline__ double as_erf_dev(double x) {
    const double a1 = 0.254829592, a2 = -0.284496736, a3 = 1.421413741;
    const double a4 = -1.453152027, a5 = 1.061405429, p = 0.3275911;
    double sign = 1.0;
    if (x < 0) sign = -1.0;
    x = fabs(x);
    const double t = 1.0 / (1.0 + p * x);
    const double y = 1.0 - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * exp(-x * x);
    return sign * y;
}

struct GTerms {
    double g, gp, gb;
};

// g(d), g'(d), g''(d): lane q of each quad evaluates the erf/exp pair of argument
// q in {A,B,C,D}; the eight values are exchanged with width-4 shuffles and every
// lane combines them in the oracle's order.  Must be called by all 32 lanes.
__device__ __forceinline__ GTerms g_terms_quad(double d, const ScoreConsts& c, double c_min, double c_max) {
    const int q = threadIdx.x & 3;
    double X;
    if (q == 0) X = d + 2 * c.r - 1 - c.mean;
    else if (q == 1) X = c_min + d + c.r - c.mean;
    else if (q == 2) X = c_max + d + c.r - c.mean;
    else X = c_min + c_max + d + 1 - c.mean;
    const double z = X / c.s2;
    const double e = c.erf_variant == BESST_ERF_LIBM ? erf(z) : as_erf_dev(z);
    const double x = exp(-(X * X) / c.v2);
    const double eA = __shfl_sync(0xffffffffu, e, 0, 4), eB = __shfl_sync(0xffffffffu, e, 1, 4);
    const double eC = __shfl_sync(0xffffffffu, e, 2, 4), eD = __shfl_sync(0xffffffffu, e, 3, 4);
    const double xA = __shfl_sync(0xffffffffu, x, 0, 4), xB = __shfl_sync(0xffffffffu, x, 1, 4);
    const double xC = __shfl_sync(0xffffffffu, x, 2, 4), xD = __shfl_sync(0xffffffffu, x, 3, 4);
    const double term1 = (c_min - c.r + 1) / 2.0 * (eC - eB);
    const double term2 = (c_min + c_max + d - c.mean + 1) / 2.0 * (eD - eC);
    const double term3 = (d + 2 * c.r - c.mean - 1) / 2.0 * (eA - eB);
    const double term4 = c.k * (xD + xA);
    const double term5 = -c.k * (xC + xB);
    GTerms t;
    t.g = term1 + term2 + term3 + term4 + term5;
    t.gp = 0.5 * (eA - eB) + 0.5 * (eD - eC);
    t.gb = (xA - xB - xC + xD) / c.gb_den;
    return t;
}

// GapEstimator: bisection on d for  mean - mean_obs = d + sd^2 g'(d)/g(d).
// `active` lets a quad idle through the loop (all 32 lanes must keep shuffling).
__device__ __forceinline__ int gap_estimator_quad(const ScoreConsts& c, double mean_obs, double c1, double c2, bool active) {
    const double obs = c.mean - mean_obs;
    const double c_min = c1 < c2 ? c1 : c2, c_max = c1 < c2 ? c2 : c1;
    double d_upper = c.d_upper0, d_lower = c.d_lower0;
    for (;;) {
        const bool more = active && (d_upper - d_lower > 1);
        if (!__any_sync(0xffffffffu, more)) break;
        const double d_ml = (d_upper + d_lower) / 2.0;
        const GTerms t = g_terms_quad(d_ml, c, c_min, c_max);
        if (more) {
            const double aofd = t.gp / t.g;
            const double func_of_d = d_ml + aofd * c.sd2;
            if (func_of_d > obs) d_upper = d_ml; else d_lower = d_ml;
        }
    }
    return (int)rint((d_upper + d_lower) / 2.0);
}

__device__ __forceinline__ double tr_sk_std_dev_quad(const ScoreConsts& c, double c1, double c2, double d) {
    const double c_min = c1 < c2 ? c1 : c2, c_max = c1 < c2 ? c2 : c1;
    const GTerms t = g_terms_quad(d, c, c_min, c_max);
    const double r1 = t.gp / t.g, r2 = t.gb / t.g;
    const double e_x = c.mean - c.sd2 * r1;
    const double e_x_square = c.sd2 + c.mean2 + c.sd4 * r2 - 2 * c.mean * c.sd2 * r1;
    const double e_o = e_x - d;
    const double e_o_square = e_x_square - 2 * d * e_x + d * d;
    const double var = e_o_square - e_o * e_o;
    if (!(var >= 0)) return 0.0;
    return sqrt(var);
}

// ---- K5: KS statistic from two device-wide radix sorts --------------------------------------------
// Links of large-large edges are compacted into "LL link space" (edge order, ll_off[e] = first slot of
// edge e).  Two 32/64-bit keys per link, (e << B) | value, are radix sorted (keys only): list 1 holds
// the observations on edge_u's scaffold, list 2 holds max(obs_v) - obs_v (CreateGraph.py:582-593).
// The sort leaves every edge's two lists ascending in place; the KS statistic
// (scipy.stats.ks_2samp(...).statistic, :595) is then a co-ranking walk over fixed-size chunks of the
// sorted arrays -- load-balanced by links, not by edges -- with a per-edge atomicMax on the bit
// pattern of the (non-negative) double.

struct KsArgs {
    const long long* row_ptr;    // [E+1] CSR
    const u32* ll_off;           // [E+1] exclusive sum of nr_links over LL edges
    const unsigned char* flags;  // [E]
    const long long* sum_u;      // [E] sum of obs_u
    const long long* obs_sum;    // [E] sum of obs_u + obs_v
    const int* max_v;            // [E]
    long long n_edges;
    long long n_ll;              // LL links
    int value_bits;
};

// edge owning LL slot j: last e with ll_off[e] <= j (zero-length non-LL entries never win a slot)
__device__ __forceinline__ long long ll_edge_of(const u32* __restrict__ ll_off, long long n_edges, u32 j) {
    long long lo = 0, hi = n_edges;   // invariant: ll_off[lo] <= j < ll_off[hi]
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(ll_off + mid) <= j) lo = mid; else hi = mid;
    }
    return lo;
}

template <typename KeyT>
__global__ void __launch_bounds__(256)
    k_score_keys(const KsArgs A, const int* __restrict__ obs_u, const int* __restrict__ obs_v, KeyT* __restrict__ key1,
                 KeyT* __restrict__ key2) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long e = warp_global; e < A.n_edges; e += n_warps) {
        if (A.ll_off[e + 1] == A.ll_off[e]) continue;   // not in this LL link space
        const long long b = A.row_ptr[e], t = A.row_ptr[e + 1];
        const long long dst = (long long)A.ll_off[e] - b;
        const int mv = A.max_v[e];
        const KeyT hi = (KeyT)((unsigned long long)e << A.value_bits);
        for (long long j = b + lane; j < t; j += 32) {
            key1[dst + j] = hi | (KeyT)(u32)__ldg(obs_u + j);
            key2[dst + j] = hi | (KeyT)(u32)(mv - __ldg(obs_v + j));   // abs(x - max_obs2), :588-590
        }
    }
}

constexpr int KS_CHUNK = 16;

// SIDE 0: evaluate at the points of list 1 (own = key1, other = key2); SIDE 1: the reverse.
template <typename KeyT, int SIDE>
__global__ void __launch_bounds__(256)
    k_ks_eval(const KsArgs A, const KeyT* __restrict__ own, const KeyT* __restrict__ other, double* ks) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long j = c * KS_CHUNK;
    if (j >= A.n_ll) return;
    const long long j_end = (j + KS_CHUNK < A.n_ll) ? j + KS_CHUNK : A.n_ll;
    const KeyT vmask = (KeyT)((1ull << A.value_bits) - 1ull);
    long long e = ll_edge_of(A.ll_off, A.n_edges, (u32)j);
    while (j < j_end) {
        // segment of edge e in LL space
        while (!(A.flags[e] & BESST_EDGE_LL) || (long long)A.ll_off[e + 1] <= j) ++e;
        const long long s0 = A.ll_off[e], s1 = A.ll_off[e + 1];
        const int n = (int)(s1 - s0);
        const long long su = A.sum_u[e];
        const long long sy = (long long)n * A.max_v[e] - (A.obs_sum[e] - su);
        const double m1 = (double)su / (double)n;   // l1_mean (:584)
        const double m2 = (double)sy / (double)n;   // l2_mean (:591)
        const double m_own = SIDE == 0 ? m1 : m2, m_other = SIDE == 0 ? m2 : m1;
        const long long stop = j_end < s1 ? j_end : s1;
        // co-rank of the first point in the other list: #{y : y - m_other <= x - m_own}
        long long q;
        {
            const double z = (double)(long long)(__ldg(own + j) & vmask) - m_own;
            long long lo = s0, hi = s1;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if ((double)(long long)(__ldg(other + mid) & vmask) - m_other <= z) lo = mid + 1; else hi = mid;
            }
            q = lo;
        }
        double dmax = 0.0;
        KeyT cur = __ldg(own + j);
        for (; j < stop; ++j) {
            const bool at_end = j + 1 >= s1;
            const KeyT nxt = at_end ? cur : __ldg(own + j + 1);
            const double z = (double)(long long)(cur & vmask) - m_own;
            while (q < s1 && (double)(long long)(__ldg(other + q) & vmask) - m_other <= z) ++q;
            if (at_end || nxt != cur) {   // last of a run of equal values: ECDF of the own list = (j + 1 - s0) / n
                const double f_own = (double)(j + 1 - s0) / (double)n;
                const double f_other = (double)(q - s0) / (double)n;
                const double diff = SIDE == 0 ? fabs(f_own - f_other) : fabs(f_other - f_own);
                if (diff > dmax) dmax = diff;
            }
            cur = nxt;
        }
        if (dmax > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(ks + e), (unsigned long long)__double_as_longlong(dmax));
    }
}

struct ScoreArgs {
    EdgeArrays E;
    long long n_edges;
    const int* scaf_len;
    ScoreConsts c;
};

// ---- K6: GapEstimator + tr_sk_std_dev + score, one quad of lanes per edge (CreateGraph.py:501-614) ----
__global__ void __launch_bounds__(128) k_edge_finalize(const ScoreArgs A) {
    const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool in_range = e < A.n_edges;
    const bool ll = in_range && (A.E.flags[e] & BESST_EDGE_LL);
    const ScoreConsts& c = A.c;
    const int n = ll ? A.E.nr[e] : 1;
    const double len1 = ll ? (double)A.scaf_len[A.E.u[e] >> 1] : 1.0, len2 = ll ? (double)A.scaf_len[A.E.v[e] >> 1] : 1.0;
    const long long obs = ll ? A.E.obs[e] : 0, obs_sq = ll ? A.E.obs_sq[e] : 0;
    const double dmax = ll ? A.E.ks[e] : 0.0;   // accumulated by k_ks_eval (initialised to +0.0)
    unsigned char flags = ll ? (A.E.flags[e] | BESST_EDGE_SCORED) : 0;
    const double mean_ = (double)obs / (double)n;                                     // :505
    const double data_observation = ((double)n * c.mean - (double)obs) / (double)n;   // :511
    const bool big = ll && (2 * c.sd < len1) && (2 * c.sd < len2);                    // :536
    const int gap_ml = gap_estimator_quad(c, mean_, len1, len2, big);
    double gap = data_observation;
    if (big) { gap = (double)gap_ml; flags |= BESST_EDGE_BIG; }
    const int gap_int = (int)gap;                                                     // :541
    const bool neg = (-gap > len1) || (-gap > len2);                                  // :542
    const double sd_ml = tr_sk_std_dev_quad(c, len1, len2, gap);
    if (!ll || (threadIdx.x & 3) != 0) return;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double score = 0.0, sd_model_out = nan;
    // the sample sd and the KS statistic are reported for every scored edge, also when the score is skipped
    // (:542-544): diagnostics, and what the lognormal scoring branch re-derives its verdict from
    double std_dev;
    if (n - 1 == 0) std_dev = 4294967296.0;                                           // :563-564
    else {
        const double q = ((double)obs_sq - (double)n * (mean_ * mean_)) / (double)(n - 1);
        if (q < 0) { std_dev = nan; flags |= BESST_EDGE_CPLX; }
        else std_dev = sqrt(q);                                                       // :561
    }
    const double ks_out = dmax, sd_obs_out = std_dev;
    if (neg) {
        flags |= BESST_EDGE_NEGGAP;
    } else {
        const double std_dev_d_eq_0 = big ? sd_ml : 4294967296.0;                    // :548-558
        const double span_score = n < 5 ? 0.0 : 1 - dmax;                             // :603-606
        double std_dev_score;
        if (std_dev_d_eq_0 == 0.0 || std_dev == 0.0 || std_dev != std_dev) std_dev_score = 0.0;
        else {
            const double x = std_dev / std_dev_d_eq_0, y = std_dev_d_eq_0 / std_dev;
            std_dev_score = y < x ? y : x;
        }
        score = (std_dev_score > 0.5 && span_score > 0.5) ? std_dev_score + span_score : 0.0;  // :614
        sd_model_out = std_dev_d_eq_0;
    }
    A.E.gap[e] = gap_int;
    A.E.score[e] = score;
    A.E.ks[e] = ks_out;
    A.E.sd_obs[e] = sd_obs_out;
    A.E.sd_model[e] = sd_model_out;
    A.E.flags[e] = flags;
}

// ---- exclusive scan of (LL ? nr_links : 0) over the edges -> ll_off[E+1] -------------------------------
constexpr int LS_THREADS = 256;
constexpr int LS_ITEMS = 8;
constexpr int LS_TILE = LS_THREADS * LS_ITEMS;

// LL link space of the edges with more than `thr` links (thr = 0: every LL edge)
__global__ void __launch_bounds__(LS_THREADS) k_ll_count(const unsigned char* __restrict__ flags, const int* __restrict__ nr,
                                                         long long n_edges, int thr, u32* block_sums) {
    __shared__ u32 s_w[LS_THREADS / 32];
    const long long base = (long long)blockIdx.x * LS_TILE;
    u32 c = 0;
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) {
        const long long e = base + i * LS_THREADS + threadIdx.x;
        if (e < n_edges && (flags[e] & BESST_EDGE_LL) && nr[e] > thr) c += (u32)nr[e];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0;
        for (int w = 0; w < LS_THREADS / 32; ++w) t += s_w[w];
        block_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(LS_THREADS) k_ll_write(const unsigned char* __restrict__ flags, const int* __restrict__ nr,
                                                         long long n_edges, int thr, const u32* __restrict__ block_sums, u32* ll_off) {
    __shared__ u32 s_w[LS_THREADS / 32];
    const long long base = (long long)blockIdx.x * LS_TILE + (long long)threadIdx.x * LS_ITEMS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 v[LS_ITEMS], c = 0;
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) {
        const long long e = base + i;
        v[i] = (e < n_edges && (flags[e] & BESST_EDGE_LL) && nr[e] > thr) ? (u32)nr[e] : 0u;
        c += v[i];
    }
    u32 incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    u32 pos = block_sums[blockIdx.x] + incl - c;
    for (int w = 0; w < warp; ++w) pos += s_w[w];
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) {
        if (base + i < n_edges) ll_off[base + i] = pos;
        pos += v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ll_off[n_edges] = block_sums[gridDim.x];
}

// ---- K5': KS statistic of the edges with at most KB_G links, one CTA per ~KB_W links ---------------------
// The per-edge observation lists are contiguous (CSR order).  A CTA takes the LL edges whose first
// link falls into its window of the compact "small LL link space", builds (local edge << B | value)
// keys for both lists in shared memory, sorts each with an in-block LSD radix sort (stable warp
// multisplit, 8-bit digits, only as many passes as the key has bits) and evaluates the two-sided
// ECDF difference -- all in one launch, one read of the observations, no global sort passes.
// Evaluation: every element finds its rank in the other list by binary search (shared memory); the
// statistic of an edge is max |i/n - k/n| over (own rank i, other rank k).  For one edge the fp64
// value is strictly increasing in |i - k| (steps of 1/n >= 2^-11 against rounding errors of 2^-52),
// so the per-edge maximum of the INTEGER |i - k| is taken first (shared-memory atomics) and only
// its maximisers evaluate scipy's expression fabs(i/n - k/n) in fp64: bit-identical, two fp64
// divisions per edge instead of per link.  Edges with more links take the device-wide path above.
constexpr int KB_THREADS = 256;
constexpr int KB_WARPS = KB_THREADS / 32;
constexpr int KB_G = 2048;            // largest edge handled here
constexpr int KB_W = 768;             // window of the small LL link space per CTA (>= the number of its edges)
constexpr int KB_CAP = 3072;          // links per CTA: < KB_W (window) + KB_G (the edge straddling its end)
static_assert(KB_W + KB_G <= KB_CAP + 1, "CTA capacity");

struct KsBlockSmem {
    u32 buf[3][KB_CAP];                      // two key lists + ping-pong; the free one holds the per-edge means afterwards
    unsigned short warp_hist[KB_WARPS][256]; // later: u32 max |i - k| per local edge
    u32 digit_start[256];
    u32 warp_sum[KB_WARPS];
    u32 seg_start[KB_W + 1];                 // first position of every local edge
};
static_assert(sizeof(double2) * KB_W <= sizeof(u32) * KB_CAP, "means fit the free key buffer");
static_assert(sizeof(u32) * KB_W <= sizeof(unsigned short) * KB_WARPS * 256, "integer maxima fit the histogram area");

// stable LSD pass over `count` <= ITEMS * 256 keys: src -> dst by digit (key >> shift) & 255
template <int ITEMS>
__device__ __forceinline__ void kb_radix_pass(KsBlockSmem& S, const u32* src, u32* dst, int count, int shift) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < KB_WARPS * 256 / 2; i += KB_THREADS) reinterpret_cast<u32*>(&S.warp_hist[0][0])[i] = 0;
    __syncthreads();
    u32 key[ITEMS];
    unsigned short rank[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int p = warp * (32 * ITEMS) + i * 32 + lane;
        const bool valid = p < count;
        const u32 vmask = __ballot_sync(0xffffffffu, valid);
        key[i] = 0; rank[i] = 0;
        if (valid) {
            key[i] = src[p];
            const u32 d = (key[i] >> shift) & 255u;
            const u32 peers = __match_any_sync(vmask, d);
            const int leader = __ffs(peers) - 1;
            u32 pre = 0;
            if (lane == leader) {
                pre = S.warp_hist[warp][d];
                S.warp_hist[warp][d] = (unsigned short)(pre + __popc(peers));
            }
            pre = __shfl_sync(vmask, pre, leader);
            rank[i] = (unsigned short)(pre + __popc(peers & lt_mask));
        }
        __syncwarp();
    }
    __syncthreads();
    {   // thread d owns digit d: prefix over warps, then exclusive scan over digits
        const int d = threadIdx.x;
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < KB_WARPS; ++w) { const u32 t = S.warp_hist[w][d]; S.warp_hist[w][d] = (unsigned short)run; run += t; }
        u32 incl = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) S.warp_sum[warp] = incl;
        __syncthreads();
        u32 wbase = 0;
#pragma unroll
        for (int w = 0; w < KB_WARPS; ++w)
            if (w < warp) wbase += S.warp_sum[w];
        S.digit_start[d] = wbase + incl - run;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int p = warp * (32 * ITEMS) + i * 32 + lane;
        if (p < count) {
            const u32 d = (key[i] >> shift) & 255u;
            dst[S.digit_start[d] + S.warp_hist[warp][d] + rank[i]] = key[i];
        }
    }
    __syncthreads();
}

struct KsBlockArgs {
    const long long* row_ptr;
    const int* nr;
    const long long* sum_u;
    const long long* obs_sum;
    const int* max_v;
    const int* obs_u;
    const int* obs_v;
    const u32* ll_edges;   // [n_small] global edge id of every small LL edge, edge order
    const u32* ll_start;   // [n_small + 1] exclusive prefix of their link counts
    long long n_small;
    long long n_links;     // ll_start[n_small]
    int value_bits;
    double* ks;
};

// rank of own[j] in the other list of its edge: #{y : y - m_other <= x - m_own}, and whether own[j] is
// the last of a run of equal values (ks_2samp evaluates the ECDFs at distinct points)
struct KbRank {
    int lid, i, k;   // local edge, own rank (1-based, ties counted), other rank; lid < 0: not an evaluation point
};

__device__ __forceinline__ KbRank kb_rank(const KsBlockSmem& S, const double2* means, const u32* own, const u32* other, int j,
                                          int count, int value_bits, int side) {
    KbRank r;
    r.lid = -1; r.i = 0; r.k = 0;
    if (j >= count) return r;
    const u32 vmask = (1u << value_bits) - 1u;
    const u32 key = own[j];
    const int lid = (int)(key >> value_bits);
    const int s0 = (int)S.seg_start[lid], s1 = (int)S.seg_start[lid + 1];
    if (j + 1 < s1 && own[j + 1] == key) return r;   // not the last of its run of ties
    const double2 m = means[lid];
    const double m_own = side == 0 ? m.x : m.y, m_other = side == 0 ? m.y : m.x;
    const double z = (double)(long long)(key & vmask) - m_own;
    int lo = s0, hi = s1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((double)(long long)(other[mid] & vmask) - m_other <= z) lo = mid + 1; else hi = mid;
    }
    r.lid = lid; r.i = j + 1 - s0; r.k = lo - s0;
    return r;
}

template <int ITEMS>
__device__ __forceinline__ void kb_sort_and_eval(KsBlockSmem& S, const KsBlockArgs& A, int count, long long k_lo, int n_local) {
    const int lane = threadIdx.x & 31;
    u32* l1 = S.buf[0];
    u32* l2 = S.buf[1];
    u32* tmp = S.buf[2];
    int key_bits = A.value_bits;
    for (int t = n_local - 1; t > 0; t >>= 1) ++key_bits;
    // list 1: l1 <-> tmp; list 2: l2 <-> whichever of the two is free afterwards
    for (int shift = 0; shift < key_bits; shift += 8) {
        kb_radix_pass<ITEMS>(S, l1, tmp, count, shift);
        u32* x = l1; l1 = tmp; tmp = x;
    }
    for (int shift = 0; shift < key_bits; shift += 8) {
        kb_radix_pass<ITEMS>(S, l2, tmp, count, shift);
        u32* x = l2; l2 = tmp; tmp = x;
    }
    // per-edge means of the two lists (CreateGraph.py:584,591) and integer maxima
    double2* means = reinterpret_cast<double2*>(tmp);
    u32* gap_max = reinterpret_cast<u32*>(&S.warp_hist[0][0]);
    for (int lid = threadIdx.x; lid < n_local; lid += KB_THREADS) {
        const long long e = __ldg(A.ll_edges + k_lo + lid);
        const int n = (int)(S.seg_start[lid + 1] - S.seg_start[lid]);
        const long long su = A.sum_u[e];
        const long long sy = (long long)n * A.max_v[e] - (A.obs_sum[e] - su);
        means[lid] = make_double2((double)su / (double)n, (double)sy / (double)n);
        gap_max[lid] = 0;
    }
    __syncthreads();
    // pass A: integer |i - k| of every evaluation point of both sides, per-edge maximum
    unsigned short gi[2][ITEMS], gk[2][ITEMS];
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const u32* own = side == 0 ? l1 : l2;
        const u32* other = side == 0 ? l2 : l1;
#pragma unroll
        for (int t = 0; t < ITEMS; ++t) {
            const int j = t * KB_THREADS + threadIdx.x;
            const KbRank r = kb_rank(S, means, own, other, j, count, A.value_bits, side);
            gi[side][t] = (unsigned short)r.i;
            gk[side][t] = (unsigned short)r.k;
            const u32 gap = (u32)(r.i > r.k ? r.i - r.k : r.k - r.i);
            // lanes hold consecutive positions: equal edges are contiguous -> one atomic per (warp, edge)
            const u32 peers = __match_any_sync(0xffffffffu, r.lid);
            const u32 best = __reduce_max_sync(peers, gap);
            if (r.lid >= 0 && lane == __ffs(peers) - 1 && best > 0) atomicMax(gap_max + r.lid, best);
        }
    }
    __syncthreads();
    // pass B: the maximisers evaluate fabs(F_own - F_other) exactly like scipy and publish the edge's statistic
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const u32* own = side == 0 ? l1 : l2;
#pragma unroll
        for (int t = 0; t < ITEMS; ++t) {
            const int j = t * KB_THREADS + threadIdx.x;
            const int i = gi[side][t], k = gk[side][t];
            if (j >= count || i == 0) continue;   // i >= 1 for every evaluation point
            const u32 gap = (u32)(i > k ? i - k : k - i);
            const int lid = (int)(own[j] >> A.value_bits);
            if (gap == 0 || gap != gap_max[lid]) continue;
            const int n = (int)(S.seg_start[lid + 1] - S.seg_start[lid]);
            const double f_own = (double)i / (double)n, f_other = (double)k / (double)n;
            const double diff = side == 0 ? fabs(f_own - f_other) : fabs(f_other - f_own);
            const long long e = __ldg(A.ll_edges + k_lo + lid);
            atomicMax(reinterpret_cast<unsigned long long*>(A.ks + e), (unsigned long long)__double_as_longlong(diff));
        }
    }
}

#ifndef BESST_KB_MIN_CTAS
#define BESST_KB_MIN_CTAS 5   // 45 KB of shared memory per CTA: five fit an SM; 0.76 -> 0.68 ms at config 3 (more CTAs to overlap the barriers)
#endif
__global__ void __launch_bounds__(KB_THREADS, BESST_KB_MIN_CTAS) k_ks_block(const KsBlockArgs A) {
    __shared__ KsBlockSmem S;
    __shared__ long long s_k[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 2) {   // first small LL edge starting at or after the window's start / end
        const u32 target = (u32)(((long long)blockIdx.x + threadIdx.x) * KB_W);
        long long lo = 0, hi = A.n_small;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (__ldg(A.ll_start + mid) < target) lo = mid + 1; else hi = mid;
        }
        s_k[threadIdx.x] = lo;
    }
    __syncthreads();
    const long long k_lo = s_k[0], k_hi = s_k[1];
    const int n_local = (int)(k_hi - k_lo);
    if (n_local <= 0) return;
    const u32 base = __ldg(A.ll_start + k_lo);
    const int count = (int)(__ldg(A.ll_start + k_hi) - base);
    for (int i = threadIdx.x; i <= n_local; i += KB_THREADS) S.seg_start[i] = __ldg(A.ll_start + k_lo + i) - base;
    __syncthreads();
    // keys of both lists, edge by edge (one warp per edge)
    for (int lid = warp; lid < n_local; lid += KB_WARPS) {
        const long long e = __ldg(A.ll_edges + k_lo + lid);
        const long long b = A.row_ptr[e];
        const int n = (int)(S.seg_start[lid + 1] - S.seg_start[lid]);
        const int mv = A.max_v[e];
        const u32 hi = (u32)lid << A.value_bits;
        const int s0 = (int)S.seg_start[lid];
        for (int k = lane; k < n; k += 32) {
            S.buf[0][s0 + k] = hi | (u32)__ldg(A.obs_u + b + k);
            S.buf[1][s0 + k] = hi | (u32)(mv - __ldg(A.obs_v + b + k));   // abs(x - max_obs2), :588-590
        }
    }
    __syncthreads();
    if (count <= 4 * KB_THREADS) kb_sort_and_eval<4>(S, A, count, k_lo, n_local);
    else if (count <= 6 * KB_THREADS) kb_sort_and_eval<6>(S, A, count, k_lo, n_local);
    else if (count <= 8 * KB_THREADS) kb_sort_and_eval<8>(S, A, count, k_lo, n_local);
    else kb_sort_and_eval<12>(S, A, count, k_lo, n_local);
}

// compact list of the LL edges with at most `thr` links: packed sums (edges << 32 | links)
__global__ void __launch_bounds__(LS_THREADS) k_llc_count(const unsigned char* __restrict__ flags, const int* __restrict__ nr,
                                                          long long n_edges, int thr, u64* block_sums) {
    __shared__ u64 s_w[LS_THREADS / 32];
    const long long base = (long long)blockIdx.x * LS_TILE;
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) {
        const long long e = base + i * LS_THREADS + threadIdx.x;
        if (e < n_edges && (flags[e] & BESST_EDGE_LL) && nr[e] <= thr) c += (1ull << 32) | (u64)(u32)nr[e];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (int w = 0; w < LS_THREADS / 32; ++w) t += s_w[w];
        block_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(LS_THREADS) k_llc_write(const unsigned char* __restrict__ flags, const int* __restrict__ nr,
                                                          long long n_edges, int thr, const u64* __restrict__ block_sums,
                                                          u32* ll_edges, u32* ll_start) {
    __shared__ u64 s_w[LS_THREADS / 32];
    const long long base = (long long)blockIdx.x * LS_TILE + (long long)threadIdx.x * LS_ITEMS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 v[LS_ITEMS], c = 0;
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) {
        const long long e = base + i;
        v[i] = (e < n_edges && (flags[e] & BESST_EDGE_LL) && nr[e] <= thr) ? ((1ull << 32) | (u64)(u32)nr[e]) : 0ull;
        c += v[i];
    }
    u64 incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    u64 pos = block_sums[blockIdx.x] + incl - c;
    for (int w = 0; w < warp; ++w) pos += s_w[w];
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) {
        if (v[i]) {
            const u32 k = (u32)(pos >> 32);
            ll_edges[k] = (u32)(base + i);
            ll_start[k] = (u32)pos;
        }
        pos += v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const u64 tot = block_sums[gridDim.x];
        ll_start[(u32)(tot >> 32)] = (u32)tot;
    }
}

// ---- batched GapEstimator + tr_sk_std_dev: one quad per item -------------------------------
__global__ void __launch_bounds__(128)
    k_gapest_batch(const ScoreConsts c, const double* __restrict__ mean_obs, const double* __restrict__ len1,
                   const double* __restrict__ len2, long long n, int* gap_out, double* sd_out) {
    const long long quad = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool active = quad < n;
    const double mo = active ? mean_obs[quad] : 0.0;
    const double l1 = active ? len1[quad] : 1.0, l2 = active ? len2[quad] : 1.0;
    const int gap = gap_estimator_quad(c, mo, l1, l2, active);
    const double sd = tr_sk_std_dev_quad(c, l1, l2, (double)gap);
    if (active && (threadIdx.x & 3) == 0) {
        gap_out[quad] = gap;
        if (sd_out) sd_out[quad] = sd;
    }
}

__global__ void __launch_bounds__(128)
    k_trsk_sd_batch(const ScoreConsts c, const double* __restrict__ gap, const double* __restrict__ len1,
                    const double* __restrict__ len2, long long n, double* sd_out) {
    const long long quad = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool active = quad < n;
    const double d = active ? gap[quad] : 0.0;
    const double l1 = active ? len1[quad] : 1.0, l2 = active ? len2[quad] : 1.0;
    const double sd = tr_sk_std_dev_quad(c, l1, l2, d);
    if (active && (threadIdx.x & 3) == 0) sd_out[quad] = sd;
}

// d + sigma^2 g'(d)/g(d) -- the left-hand side of the ML equation (funcDGeneral) -- for a batch of d:
// what mathstats' PreCalcMLvaluesOfdLongContigs tabulates (MakeScaffolds.py:68)
__global__ void __launch_bounds__(128)
    k_func_of_d_batch(const ScoreConsts c, const double* __restrict__ d_in, const double* __restrict__ len1,
                      const double* __restrict__ len2, long long n, double* out) {
    const long long quad = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool active = quad < n;
    const double d = active ? d_in[quad] : 0.0;
    const double l1 = active ? len1[quad] : 1.0, l2 = active ? len2[quad] : 1.0;
    const double c_min = l1 < l2 ? l1 : l2, c_max = l1 < l2 ? l2 : l1;
    const GTerms t = g_terms_quad(d, c, c_min, c_max);
    const double aofd = t.gp / t.g;
    if (active && (threadIdx.x & 3) == 0) out[quad] = d + aofd * c.sd2;
}

// ---- lognormal GapEstimator (mathstats.log_normal_param_est, restated from the model: see the oracle) ---------
// one warp per edge: the lanes share the pass over the edge's observations (32 strided partial sums combined by
// a butterfly -- the summation order the oracle reproduces), g(d) in closed form from the partial moments of the
// lognormal, integer ternary search for the maximum of the log-likelihood
__device__ __forceinline__ double ln_Phi_dev(double z) { return 0.5 * erfc(-z / sqrt(2.0)); }
__device__ __forceinline__ double ln_F0_dev(double x, double mu, double sigma) { return x > 0 ? ln_Phi_dev((log(x) - mu) / sigma) : 0.0; }
__device__ __forceinline__ double ln_F1_dev(double x, double mu, double sigma) {
    return x > 0 ? exp(mu + sigma * sigma / 2.0) * ln_Phi_dev((log(x) - mu - sigma * sigma) / sigma) : 0.0;
}
__device__ __forceinline__ double ln_g_dev(double d, double mu, double sigma, double c_min, double c_max, double r) {
    const double A = d + 2 * r - 1, B = d + c_min + r, C = d + c_max + r, D = d + c_min + c_max + 1;
    const double f0A = ln_F0_dev(A, mu, sigma), f0B = ln_F0_dev(B, mu, sigma), f0C = ln_F0_dev(C, mu, sigma), f0D = ln_F0_dev(D, mu, sigma);
    const double f1A = ln_F1_dev(A, mu, sigma), f1B = ln_F1_dev(B, mu, sigma), f1C = ln_F1_dev(C, mu, sigma), f1D = ln_F1_dev(D, mu, sigma);
    const double piece1 = -(d + 2 * r - 1) * (f0B - f0A) + (f1B - f1A);
    const double piece2 = (c_min - r + 1) * (f0C - f0B);
    const double piece3 = (d + c_min + c_max + 1) * (f0D - f0C) - (f1D - f1C);
    return piece1 + piece2 + piece3;
}
__device__ __forceinline__ double ln_loglik_warp(long long d, double mu, double sigma, const int* __restrict__ samples, long long n,
                                                 double c_min, double c_max, double r, int lane) {
    const double g = ln_g_dev((double)d, mu, sigma, c_min, c_max, r);
    if (!(g > 0)) return -__longlong_as_double(0x7ff0000000000000ll);   // -inf
    double partial = 0.0;
    const double v2 = 2.0 * sigma * sigma;
    for (long long k = lane; k < n; k += 32) {
        const double lx = log((double)((long long)__ldg(samples + k) + d));
        partial += -lx - (lx - mu) * (lx - mu) / v2;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) partial += __shfl_xor_sync(0xffffffffu, partial, off);
    return partial - (double)n * log(g);
}

__global__ void __launch_bounds__(256)
    k_gapest_lognormal(double mu, double sigma, double r, const int* __restrict__ samples, const long long* __restrict__ row_ptr,
                       const double* __restrict__ len1, const double* __restrict__ len2, long long n_edges, int* gap_out) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long e = warp_global; e < n_edges; e += n_warps) {
        const long long b = row_ptr[e], n = row_ptr[e + 1] - b;
        const int* s = samples + b;
        const double c1 = len1[e], c2 = len2[e];
        const double c_min = c1 < c2 ? c1 : c2, c_max = c1 < c2 ? c2 : c1;
        int o_min = 2147483647;
        for (long long k = lane; k < n; k += 32) { const int v = __ldg(s + k); o_min = v < o_min ? v : o_min; }
        o_min = __reduce_min_sync(0xffffffffu, o_min);
        const double mean_x = exp(mu + sigma * sigma / 2.0);
        const double sd_x = sqrt((exp(sigma * sigma) - 1.0) * exp(2.0 * mu + sigma * sigma));
        long long lo = (long long)(-c_min);
        const long long alt = 1 - (long long)o_min;
        if (alt > lo) lo = alt;
        long long hi = (long long)(mean_x + 4.0 * sd_x);
        long long best = lo;
        if (hi > lo && n > 0) {
            while (hi - lo > 2) {
                const long long third = (hi - lo) / 3, m1 = lo + third, m2 = hi - third;
                const double a = ln_loglik_warp(m1, mu, sigma, s, n, c_min, c_max, r, lane);
                const double bb = ln_loglik_warp(m2, mu, sigma, s, n, c_min, c_max, r, lane);
                if (a < bb) lo = m1 + 1; else hi = m2 - 1;
            }
            best = lo;
            double best_l = ln_loglik_warp(lo, mu, sigma, s, n, c_min, c_max, r, lane);
            for (long long d = lo + 1; d <= hi; ++d) {
                const double v = ln_loglik_warp(d, mu, sigma, s, n, c_min, c_max, r, lane);
                if (v > best_l) { best = d; best_l = v; }
            }
        }
        if (lane == 0) gap_out[e] = (int)best;
    }
}

int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 32 && (max_value >> b)) ++b;
    return b;
}

}  // namespace

ScoreConsts besst_score_consts(const besst_lib_params& p) {
    ScoreConsts c;
    c.mean = p.mean_ins_size; c.sd = p.std_dev_ins_size; c.r = p.read_len;
    c.s2 = pow(2.0, 0.5) * c.sd;
    c.v2 = 2 * pow(c.sd, 2.0);
    c.k = c.sd / pow(2 * M_PI, 0.5);
    c.gb_den = pow(2 * M_PI, 0.5) * c.sd;
    c.sd2 = pow(c.sd, 2.0); c.sd4 = pow(c.sd, 4.0); c.mean2 = pow(c.mean, 2.0);
    c.d_upper0 = (double)(int64_t)(c.mean + 2 * c.sd - 2 * c.r);
    c.d_lower0 = (double)(int64_t)(-4 * c.sd);
    c.erf_variant = p.erf_variant;
    return c;
}

int besst_launch_gapest(besst_ctx* ctx, const besst_lib_params& p, const double* d_mean_obs, const double* d_len1,
                        const double* d_len2, int64_t n, int32_t* d_gap, double* d_sd) {
    if (n == 0) return BESST_OK;
    const ScoreConsts c = besst_score_consts(p);
    const long long threads = n * 4;
    const int grid = (int)((threads + 127) / 128);
    { KTimer kt(ctx, BESST_K_GAPEST); k_gapest_batch<<<grid, 128, 0, ctx->stream>>>(c, d_mean_obs, d_len1, d_len2, n, d_gap, d_sd); }
    BESST_CUDA_TRY(ctx, cudaGetLastError());
    return BESST_OK;
}

int besst_launch_gapest_lognormal(besst_ctx* ctx, double mu, double sigma, double r, const int32_t* d_samples, const int64_t* d_row_ptr,
                                  const double* d_len1, const double* d_len2, int64_t n, int32_t* d_gap) {
    if (n == 0) return BESST_OK;
    long long grid = (n * 32 + 255) / 256;
    if (grid > (long long)ctx->sm_count * 16) grid = (long long)ctx->sm_count * 16;
    { KTimer kt(ctx, BESST_K_GAPEST);
      k_gapest_lognormal<<<(unsigned)grid, 256, 0, ctx->stream>>>(mu, sigma, r, d_samples, reinterpret_cast<const long long*>(d_row_ptr), d_len1, d_len2, n, d_gap); }
    BESST_CUDA_TRY(ctx, cudaGetLastError());
    return BESST_OK;
}

int besst_launch_func_of_d(besst_ctx* ctx, const besst_lib_params& p, const double* d_d, const double* d_len1,
                           const double* d_len2, int64_t n, double* d_out) {
    if (n == 0) return BESST_OK;
    const ScoreConsts c = besst_score_consts(p);
    const int grid = (int)((n * 4 + 127) / 128);
    { KTimer kt(ctx, BESST_K_GAPEST); k_func_of_d_batch<<<grid, 128, 0, ctx->stream>>>(c, d_d, d_len1, d_len2, n, d_out); }
    BESST_CUDA_TRY(ctx, cudaGetLastError());
    return BESST_OK;
}

int besst_launch_trsk_sd(besst_ctx* ctx, const besst_lib_params& p, const double* d_gap, const double* d_len1,
                         const double* d_len2, int64_t n, double* d_sd) {
    if (n == 0) return BESST_OK;
    const ScoreConsts c = besst_score_consts(p);
    const long long threads = n * 4;
    const int grid = (int)((threads + 127) / 128);
    { KTimer kt(ctx, BESST_K_GAPEST); k_trsk_sd_batch<<<grid, 128, 0, ctx->stream>>>(c, d_gap, d_len1, d_len2, n, d_sd); }
    BESST_CUDA_TRY(ctx, cudaGetLastError());
    return BESST_OK;
}

// k_group_blocks over a device tuple array: grouped observations in ctx->grouped, run descriptors in
// ctx->run_key[0] / run_val[0] / run_start / run_cnt / run_first.  *overflow: the stream has no local order
// (a block with more than GB_DMAX edges, or more runs than n/8) -- use the radix bucket
int besst_group_tuples(besst_ctx* ctx, const besst_link_tuple* d_tuples, int64_t n, int bv, int block_bits, int64_t* n_runs,
                       int* overflow) {
    if (n_runs) { *n_runs = 0; *overflow = 0; }
    const size_t nz = (size_t)(n > 0 ? n : 1);
    const int64_t n_gblocks = (n + GB_TILE - 1) / GB_TILE;
    const int64_t run_cap64 = std::max<int64_t>(n / 8, 1 << 16);
    const u32 run_cap = (u32)std::min<int64_t>(run_cap64, 0x7fffffff);
    BESST_CUDA_TRY(ctx, ctx->grouped.ensure(8 * nz));
    for (int k = 0; k < 2; ++k) {
        BESST_CUDA_TRY(ctx, ctx->run_key[k].ensure(8 * (size_t)run_cap));
        BESST_CUDA_TRY(ctx, ctx->run_val[k].ensure(4 * (size_t)run_cap));
    }
    BESST_CUDA_TRY(ctx, ctx->run_start.ensure(4 * (size_t)run_cap)); BESST_CUDA_TRY(ctx, ctx->run_cnt.ensure(4 * (size_t)run_cap));
    BESST_CUDA_TRY(ctx, ctx->run_first.ensure(4 * (size_t)run_cap)); BESST_CUDA_TRY(ctx, ctx->run_off.ensure(4 * (size_t)run_cap));
    BESST_CUDA_TRY(ctx, ctx->run_src.ensure(4 * (size_t)run_cap)); BESST_CUDA_TRY(ctx, ctx->run_len.ensure(4 * (size_t)run_cap));
    BESST_CUDA_TRY(ctx, ctx->run_state.ensure(512));
    u32* gstate = ctx->run_state.as<u32>();
    BESST_CUDA_TRY(ctx, cudaMemsetAsync(gstate, 0, 64, ctx->stream));
    if (n == 0) return BESST_OK;
    if (!ctx->attr_group_done) {   // per ctx (= per device): function attributes are set on the current device
        cudaFuncSetAttribute(k_group_blocks<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GroupSmem));
        cudaFuncSetAttribute(k_group_blocks<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GroupSmem));
        ctx->attr_group_done = true;
    }
    {
        KTimer kt(ctx, BESST_K_GROUP);
        if (d_tuples)
            k_group_blocks<false><<<(unsigned)n_gblocks, GB_THREADS, sizeof(GroupSmem), ctx->stream>>>(
                d_tuples, n, bv, block_bits, ctx->grouped.as<int2>(), ctx->run_key[0].as<u64>(), ctx->run_val[0].as<u32>(),
                ctx->run_start.as<u32>(), ctx->run_cnt.as<u32>(), ctx->run_first.as<u32>(), gstate, run_cap, nullptr, 0, nullptr);
        else   // the last extraction, in place
            k_group_blocks<true><<<(unsigned)n_gblocks, GB_THREADS, sizeof(GroupSmem), ctx->stream>>>(
                ctx->scratch_tuples.as<besst_link_tuple>(), n, bv, block_bits, ctx->grouped.as<int2>(), ctx->run_key[0].as<u64>(),
                ctx->run_val[0].as<u32>(), ctx->run_start.as<u32>(), ctx->run_cnt.as<u32>(), ctx->run_first.as<u32>(), gstate, run_cap,
                ctx->tile_state.as<u64>(), ctx->n_rec_tiles, ctx->block_tile0.as<u32>());
    }
    BESST_CUDA_TRY(ctx, cudaGetLastError());
    if (!n_runs) return BESST_OK;   // queued only: the caller reads run_state[0..1] together with its other sizes
    u32* const hs = reinterpret_cast<u32*>(ctx->host_scalars());
    if (!hs) { ctx->err = "pinned host scratch allocation failed"; return BESST_E_NOMEM; }
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(hs, gstate, 8, cudaMemcpyDeviceToHost, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *n_runs = hs[0];
    *overflow = (hs[1] != 0 || (int64_t)hs[0] > (int64_t)run_cap) ? 1 : 0;
    return BESST_OK;
}

// ---- run-level multi-GPU exchange: route / pack on the sender, import on the receiver ------------------
constexpr int RX_MAX_WORLD = 16;
// where this rank's segment starts inside every destination's observation / descriptor buffer: local
// addresses (NCCL all-to-all afterwards) or peer-mapped addresses of the destination GPU (NVLink stores)
struct PackDst { void* obs[RX_MAX_WORLD]; besst_run_desc* desc[RX_MAX_WORLD]; };   // obs: int2 or (PACK16) u32 per link

// state: [0..15] links per destination, [16..31] runs per destination
// R_dev != nullptr: the number of runs is still on the device (run_state[0], capped at run_cap by the caller's check later)
__global__ void __launch_bounds__(256) k_runs_route_count(const u64* __restrict__ run_key, const u32* __restrict__ run_cnt, long long R,
                                                          const u32* __restrict__ R_dev, long long run_cap, int block_bits, int bv, int world, u32* state) {
    __shared__ u32 s_l[RX_MAX_WORLD], s_r[RX_MAX_WORLD];
    if (threadIdx.x < RX_MAX_WORLD) { s_l[threadIdx.x] = 0; s_r[threadIdx.x] = 0; }
    __syncthreads();
    if (R_dev) { R = (long long)*R_dev; if (R > run_cap) R = run_cap; }
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (long long)gridDim.x * blockDim.x) {
        const u64 key = run_key[r] >> block_bits;
        const u32 d = besst_edge_dest((u32)(key >> bv), (u32)(key & ((1ull << bv) - 1ull)), world);
        atomicAdd(&s_l[d], run_cnt[r]);
        atomicAdd(&s_r[d], 1u);
    }
    __syncthreads();
    if (threadIdx.x < world) {
        if (s_l[threadIdx.x]) atomicAdd(state + threadIdx.x, s_l[threadIdx.x]);
        if (s_r[threadIdx.x]) atomicAdd(state + RX_MAX_WORLD + threadIdx.x, s_r[threadIdx.x]);
    }
}

// one warp per run: reserve space in its destination segment, copy the observations, write the descriptor
template <bool PACK16>
__global__ void __launch_bounds__(256) k_runs_pack(const u64* __restrict__ run_key, const u32* __restrict__ run_start,
                                                   const u32* __restrict__ run_cnt, const u32* __restrict__ run_first, long long R,
                                                   int block_bits, int bv, int world, const int2* __restrict__ grouped,
                                                   const PackDst D, u32* cursors) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp_global; r < R; r += n_warps) {
        const u64 word = __ldg(run_key + r);
        const u64 key = word >> block_bits;
        const u32 u = (u32)(key >> bv), v = (u32)(key & ((1ull << bv) - 1ull));
        const u32 d = besst_edge_dest(u, v, world);
        const u32 cnt = __ldg(run_cnt + r), src = __ldg(run_start + r);
        u32 a = 0, slot = 0;
        if (lane == 0) { a = atomicAdd(cursors + d, cnt); slot = atomicAdd(cursors + RX_MAX_WORLD + d, 1u); }
        a = __shfl_sync(0xffffffffu, a, 0);
        if (PACK16) {
            u32* dst = reinterpret_cast<u32*>(D.obs[d]) + a;
            for (u32 k = lane; k < cnt; k += 32) {
                const int2 o = __ldg(grouped + src + k);
                dst[k] = (u32)o.x | ((u32)o.y << 16);
            }
        } else {
            int2* dst = reinterpret_cast<int2*>(D.obs[d]) + a;
            for (u32 k = lane; k < cnt; k += 32) dst[k] = __ldg(grouped + src + k);
        }
        if (lane == 0) {
            besst_run_desc ds;
            ds.u = u; ds.v = v; ds.count = cnt; ds.first = __ldg(run_first + r); ds.offset = a;
            ds.block = (u32)(word & ((1ull << block_bits) - 1ull));
            D.desc[d][slot] = ds;
        }
    }
}

struct ImportBases { u32 run_end[RX_MAX_WORLD]; u32 link_base[RX_MAX_WORLD]; u32 first_base[RX_MAX_WORLD]; };

// received descriptors (source-major) -> the sort key (u, v, source, block) and absolute positions
__global__ void __launch_bounds__(256) k_runs_import(const besst_run_desc* __restrict__ desc, long long R, int world, int bv,
                                                     int block_bits, int low_bits, const ImportBases B, u64* run_key, u32* run_val,
                                                     u32* run_start, u32* run_cnt, u32* run_first) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int src = 0;
    while (src + 1 < world && (u32)r >= B.run_end[src]) ++src;
    const besst_run_desc ds = desc[r];
    run_key[r] = (((((u64)ds.u << bv) | ds.v)) << low_bits) | ((u64)src << block_bits) | (u64)ds.block;
    run_val[r] = (u32)r;
    run_start[r] = B.link_base[src] + ds.offset;
    run_cnt[r] = ds.count;
    run_first[r] = B.first_base[src] + ds.first;
}

// the route counts of runs whose number is still on the device (queued behind k_group_blocks, no host read)
int besst_launch_runs_route_async(besst_ctx* ctx, int world, int block_bits, int64_t run_cap) {
    const int bv = bits_for((uint64_t)(2 * ctx->n_scaffolds > 0 ? 2 * ctx->n_scaffolds - 1 : 1));
    u32* state = ctx->run_state.as<u32>() + 16;   // [16..47]: route counts, [48..79]: pack cursors
    BESST_CUDA_TRY(ctx, cudaMemsetAsync(state, 0, 4 * 4 * RX_MAX_WORLD, ctx->stream));
    { KTimer kt(ctx, BESST_K_PARTITION);
      k_runs_route_count<<<(unsigned)(ctx->sm_count * 8), 256, 0, ctx->stream>>>(ctx->run_key[0].as<u64>(), ctx->run_cnt.as<u32>(), 0,
                                                                                 ctx->run_state.as<u32>(), run_cap, block_bits, bv, world, state); }
    BESST_CUDA_TRY(ctx, cudaGetLastError());
    return BESST_OK;
}

int besst_launch_runs_route(besst_ctx* ctx, int world, int64_t* link_counts, int64_t* run_counts) {
    for (int d = 0; d < world; ++d) link_counts[d] = run_counts[d] = 0;
    const int64_t R = ctx->n_runs;
    if (R == 0) return BESST_OK;
    const int bv = bits_for((uint64_t)(2 * ctx->n_scaffolds > 0 ? 2 * ctx->n_scaffolds - 1 : 1));
    u32* state = ctx->run_state.as<u32>() + 16;   // [16..47]: route counts, [48..79]: pack cursors
    BESST_CUDA_TRY(ctx, cudaMemsetAsync(state, 0, 4 * 4 * RX_MAX_WORLD, ctx->stream));
    { KTimer kt(ctx, BESST_K_PARTITION);
      long long rgrid = (R + 255) / 256;
      if (rgrid > (long long)ctx->sm_count * 8) rgrid = (long long)ctx->sm_count * 8;
      k_runs_route_count<<<(unsigned)rgrid, 256, 0, ctx->stream>>>(ctx->run_key[0].as<u64>(), ctx->run_cnt.as<u32>(), R, nullptr, R,
                                                                   ctx->run_block_bits, bv, world, state); }
    BESST_CUDA_TRY(ctx, cudaGetLastError());
    u32 h[2 * RX_MAX_WORLD];
    BESST_CUDA_TRY(ctx, cudaMemcpyAsync(h, state, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int d = 0; d < world; ++d) { link_counts[d] = h[d]; run_counts[d] = h[RX_MAX_WORLD + d]; }
    return BESST_OK;
}

// obs_ptrs / desc_ptrs == nullptr: one local destination-major buffer pair (out_obs, out_desc)
int besst_launch_runs_pack(besst_ctx* ctx, int world, int32_t* out_obs, besst_run_desc* out_desc, int32_t* const* obs_ptrs,
                           besst_run_desc* const* desc_ptrs) {
    const int64_t R = ctx->n_runs;
    if (R == 0) return BESST_OK;
    const int bv = bits_for((uint64_t)(2 * ctx->n_scaffolds > 0 ? 2 * ctx->n_scaffolds - 1 : 1));
    u32* state = ctx->run_state.as<u32>() + 16;
    const bool pack16 = besst_obs_bytes(ctx->extract_params) == 4;
    PackDst D;
    if (obs_ptrs) {
        for (int d = 0; d < RX_MAX_WORLD; ++d) {
            D.obs[d] = d < world ? static_cast<void*>(obs_ptrs[d]) : nullptr;
            D.desc[d] = d < world ? desc_ptrs[d] : nullptr;
        }
    } else {
        u32 h[2 * RX_MAX_WORLD];
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(h, state, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        u32 lb = 0, rb = 0;
        for (int d = 0; d < RX_MAX_WORLD; ++d) {
            D.obs[d] = reinterpret_cast<char*>(out_obs) + (size_t)lb * (pack16 ? 4 : 8); D.desc[d] = out_desc + rb;
            if (d < world) { lb += h[d]; rb += h[RX_MAX_WORLD + d]; }
        }
    }
    u32* cursors = state + 2 * RX_MAX_WORLD;
    long long grid = (R * 32 + 255) / 256;
    if (grid > (long long)ctx->sm_count * 16) grid = (long long)ctx->sm_count * 16;
    { KTimer kt(ctx, BESST_K_PARTITION);
      if (pack16)
          k_runs_pack<true><<<(unsigned)grid, 256, 0, ctx->stream>>>(ctx->run_key[0].as<u64>(), ctx->run_start.as<u32>(), ctx->run_cnt.as<u32>(),
                                                                     ctx->run_first.as<u32>(), R, ctx->run_block_bits, bv, world,
                                                                     ctx->grouped.as<int2>(), D, cursors);
      else
          k_runs_pack<false><<<(unsigned)grid, 256, 0, ctx->stream>>>(ctx->run_key[0].as<u64>(), ctx->run_start.as<u32>(), ctx->run_cnt.as<u32>(),
                                                                      ctx->run_first.as<u32>(), R, ctx->run_block_bits, bv, world,
                                                                      ctx->grouped.as<int2>(), D, cursors); }
    BESST_CUDA_TRY(ctx, cudaGetLastError());
    return BESST_OK;
}

int besst_launch_runs_import(besst_ctx* ctx, const besst_run_desc* desc, int64_t n_runs, int world, int block_bits,
                             const int64_t* src_run_counts, const int64_t* src_link_counts, const int64_t* src_first_base,
                             int* low_bits) {
    const int bv = bits_for((uint64_t)(2 * ctx->n_scaffolds > 0 ? 2 * ctx->n_scaffolds - 1 : 1));
    const int src_bits = bits_for((uint64_t)(world > 1 ? world - 1 : 1));
    *low_bits = src_bits + block_bits;
    if (2 * bv + *low_bits > 64) { ctx->err = "runs_to_graph: sort key wider than 64 bits"; return BESST_E_INVALID; }
    const size_t cap = (size_t)(n_runs > 0 ? n_runs : 1);
    for (int k = 0; k < 2; ++k) {
        BESST_CUDA_TRY(ctx, ctx->run_key[k].ensure(8 * cap));
        BESST_CUDA_TRY(ctx, ctx->run_val[k].ensure(4 * cap));
    }
    BESST_CUDA_TRY(ctx, ctx->run_start.ensure(4 * cap)); BESST_CUDA_TRY(ctx, ctx->run_cnt.ensure(4 * cap));
    BESST_CUDA_TRY(ctx, ctx->run_first.ensure(4 * cap));
    if (n_runs == 0) return BESST_OK;
    ImportBases B;
    int64_t re = 0, lb = 0;
    for (int s = 0; s < RX_MAX_WORLD; ++s) {
        B.link_base[s] = (u32)lb; B.first_base[s] = s < world ? (u32)src_first_base[s] : 0u;
        if (s < world) { re += src_run_counts[s]; lb += src_link_counts[s]; }
        B.run_end[s] = (u32)re;
    }
    { KTimer kt(ctx, BESST_K_PARTITION);
      k_runs_import<<<(unsigned)((n_runs + 255) / 256), 256, 0, ctx->stream>>>(desc, n_runs, world, bv, block_bits, *low_bits, B,
                                                                               ctx->run_key[0].as<u64>(), ctx->run_val[0].as<u32>(),
                                                                               ctx->run_start.as<u32>(), ctx->run_cnt.as<u32>(), ctx->run_first.as<u32>()); }
    BESST_CUDA_TRY(ctx, cudaGetLastError());
    return BESST_OK;
}

// runs != nullptr: the run descriptors (ctx->run_key[0] / run_val[0] / run_start / run_cnt / run_first) and the
// grouped observations were prepared by the caller (multi-GPU import) instead of k_group_blocks
static int launch_graph_impl(besst_ctx* ctx, const besst_lib_params& p, const besst_link_tuple* d_tuples, int64_t n,
                             const uint64_t* d_fishy, int64_t n_fishy, const BesstRunInput* runs) {
    ctx->have_graph = false;
    ctx->n_links = n;
    ctx->n_edges = 0;
    ctx->n_ll_links = 0;
    ctx->sweep_kernel_id = BESST_K_RADIX_SWEEP;
    ctx->last_params = p;
    const int bv = bits_for((uint64_t)(2 * ctx->n_scaffolds > 0 ? 2 * ctx->n_scaffolds - 1 : 1));
    const size_t nz = (size_t)(n > 0 ? n : 1);

    // fishy pairs: rekey, sort (keys only)
    int rc;
    const u64* fishy_sorted = nullptr;
    if (n_fishy > 0) {
        BESST_CUDA_TRY(ctx, ctx->fishy_sorted.ensure(8 * (size_t)n_fishy));
        BESST_CUDA_TRY(ctx, ctx->fishy_tmp.ensure(8 * (size_t)n_fishy));
        { KTimer kt(ctx, BESST_K_FISHY); k_fishy_rekey<<<(int)((n_fishy + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const u64*>(d_fishy), ctx->fishy_sorted.as<u64>(), n_fishy, bv); }
        int fb = 0;
        ctx->sweep_kernel_id = BESST_K_FISHY;   // profiling: keep the small fishy-key sort out of k_radix_sweep
        rc = besst_radix_sort_keys(ctx, ctx->fishy_sorted.as<uint64_t>(), ctx->fishy_tmp.as<uint64_t>(), n_fishy, 2 * bv, &fb);
        ctx->sweep_kernel_id = BESST_K_RADIX_SWEEP;
        if (rc) return rc;
        fishy_sorted = fb ? ctx->fishy_tmp.as<u64>() : ctx->fishy_sorted.as<u64>();
    }

    EdgeArrays EA;
    int64_t E = 0;
    auto alloc_edges = [&](int64_t n_edges) -> int {
        E = n_edges;
        ctx->n_edges = E;
        const size_t Ez = (size_t)(E > 0 ? E : 1);
        BESST_CUDA_TRY(ctx, ctx->e_u.ensure(4 * Ez)); BESST_CUDA_TRY(ctx, ctx->e_v.ensure(4 * Ez));
        BESST_CUDA_TRY(ctx, ctx->e_nr.ensure(4 * Ez)); BESST_CUDA_TRY(ctx, ctx->e_obs.ensure(8 * Ez));
        BESST_CUDA_TRY(ctx, ctx->e_obs_sq.ensure(8 * Ez)); BESST_CUDA_TRY(ctx, ctx->e_first.ensure(8 * Ez));
        BESST_CUDA_TRY(ctx, ctx->e_row_ptr.ensure(8 * (Ez + 1))); BESST_CUDA_TRY(ctx, ctx->e_gap.ensure(4 * Ez));
        BESST_CUDA_TRY(ctx, ctx->e_score.ensure(8 * Ez)); BESST_CUDA_TRY(ctx, ctx->e_ks.ensure(8 * Ez));
        BESST_CUDA_TRY(ctx, ctx->e_sd_obs.ensure(8 * Ez)); BESST_CUDA_TRY(ctx, ctx->e_sd_model.ensure(8 * Ez));
        BESST_CUDA_TRY(ctx, ctx->e_fishy.ensure(4 * Ez)); BESST_CUDA_TRY(ctx, ctx->e_flags.ensure(Ez));
        BESST_CUDA_TRY(ctx, ctx->l_obs_u.ensure(4 * nz)); BESST_CUDA_TRY(ctx, ctx->l_obs_v.ensure(4 * nz));
        BESST_CUDA_TRY(ctx, ctx->e_sum_u.ensure(8 * Ez)); BESST_CUDA_TRY(ctx, ctx->e_max_v.ensure(4 * Ez));
        EA.u = ctx->e_u.as<u32>(); EA.v = ctx->e_v.as<u32>(); EA.nr = ctx->e_nr.as<int>();
        EA.obs = ctx->e_obs.as<long long>(); EA.obs_sq = ctx->e_obs_sq.as<long long>(); EA.first = ctx->e_first.as<long long>();
        EA.row_ptr = ctx->e_row_ptr.as<long long>(); EA.gap = ctx->e_gap.as<int>(); EA.score = ctx->e_score.as<double>();
        EA.ks = ctx->e_ks.as<double>(); EA.sd_obs = ctx->e_sd_obs.as<double>(); EA.sd_model = ctx->e_sd_model.as<double>();
        EA.fishy = ctx->e_fishy.as<int>(); EA.flags = ctx->e_flags.as<unsigned char>();
        EA.obs_u = ctx->l_obs_u.as<int>(); EA.obs_v = ctx->l_obs_v.as<int>();
        EA.sum_u = ctx->e_sum_u.as<long long>(); EA.max_v = ctx->e_max_v.as<int>();
        return BESST_OK;
    };
    auto finish_empty = [&]() -> int {
        const long long zero = 0;
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(EA.row_ptr, &zero, 8, cudaMemcpyHostToDevice, ctx->stream));
        BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        besst_mark(ctx); besst_mark(ctx); besst_mark(ctx);
        ctx->have_graph = true;
        return BESST_OK;
    };
    int edge_grid_max = ctx->sm_count * 32;

    // ---- K3'/K4': run-merge bucket (default) ------------------------------------------------------
    // BESST_BUCKET=radix forces the device-wide radix sort below (A/B measurements, tests of the fallback)
    const char* bucket_env = getenv("BESST_BUCKET");
    const bool force_radix = bucket_env && bucket_env[0] == 'r';
    bool done = false;
    bool have_runs = false;
    const int2* grouped_ptr = nullptr;
    int64_t R = 0;
    int low_bits = 0, run_key_bits = 0;
    const int64_t n_gblocks = (n + GB_TILE - 1) / GB_TILE;
    const int block_bits = bits_for((uint64_t)(n_gblocks > 1 ? n_gblocks - 1 : 1));
    if (runs) {
        have_runs = true;
        grouped_ptr = runs->grouped; R = runs->n_runs; low_bits = runs->low_bits; run_key_bits = 2 * bv + runs->low_bits;
        BESST_CUDA_TRY(ctx, ctx->run_off.ensure(4 * (size_t)(R + 1))); BESST_CUDA_TRY(ctx, ctx->run_src.ensure(4 * (size_t)(R + 1)));
        BESST_CUDA_TRY(ctx, ctx->run_len.ensure(4 * (size_t)(R + 1)));
    } else if (!force_radix && n > 0 && 2 * bv + block_bits <= 64) {
        int overflow = 0;
        rc = besst_group_tuples(ctx, d_tuples, n, bv, block_bits, &R, &overflow);
        if (rc) return rc;
        if (!overflow) { have_runs = true; grouped_ptr = ctx->grouped.as<int2>(); low_bits = block_bits; run_key_bits = 2 * bv + block_bits; }
    }
    if (have_runs && n > 0) {
        {
            int in_b = 0;
            rc = besst_radix_sort_pairs(ctx, ctx->run_key[0].as<uint64_t>(), ctx->run_key[1].as<uint64_t>(), ctx->run_val[0].as<uint32_t>(),
                                        ctx->run_val[1].as<uint32_t>(), R, run_key_bits, &in_b);
            if (rc) return rc;
            const u64* rkeys = ctx->run_key[in_b].as<u64>();
            const u32* rvals = ctx->run_val[in_b].as<u32>();
            besst_mark(ctx);
            const int rblocks = (int)((R + RS_TILE - 1) / RS_TILE);
            BESST_CUDA_TRY(ctx, ctx->heads.ensure(8 * (size_t)(rblocks + 2)));
            u64* bsums = ctx->heads.as<u64>();
            { KTimer kt(ctx, BESST_K_RUNS); k_run_count<<<rblocks, RS_T, 0, ctx->stream>>>(rkeys, rvals, ctx->run_cnt.as<u32>(), R, low_bits, bsums); }
            { KTimer kt(ctx, BESST_K_RUNS); k_scan_blocks64<<<1, 1024, 0, ctx->stream>>>(bsums, rblocks); }
            u64* const hp = reinterpret_cast<u64*>(ctx->host_scalars());
            if (!hp) { ctx->err = "pinned host scratch allocation failed"; return BESST_E_NOMEM; }
            BESST_CUDA_TRY(ctx, cudaMemcpyAsync(hp, bsums + rblocks, 8, cudaMemcpyDeviceToHost, ctx->stream));
            BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            const u64 tot = hp[0];
            rc = alloc_edges((int64_t)(tot >> 32));
            if (rc) return rc;
            if ((int64_t)(tot & 0xffffffffull) != n) { ctx->err = "run-merge bucket: link count mismatch"; return BESST_E_STATE; }
            BESST_CUDA_TRY(ctx, ctx->edge_run_ptr.ensure(4 * (size_t)(E + 2)));
            { KTimer kt(ctx, BESST_K_RUNS);
              k_run_write<<<rblocks, RS_T, 0, ctx->stream>>>(rkeys, rvals, ctx->run_cnt.as<u32>(), ctx->run_start.as<u32>(), ctx->run_first.as<u32>(), R,
                                                            low_bits, bv, bsums, ctx->run_off.as<u32>(), ctx->run_src.as<u32>(), ctx->run_len.as<u32>(),
                                                            EA.row_ptr, ctx->edge_run_ptr.as<u32>(), EA.u, EA.v, EA.first, n); }
            besst_mark(ctx);
            {
                int grid = (int)((E * 32 + 255) / 256);
                if (grid > edge_grid_max) grid = edge_grid_max;
                KTimer kt(ctx, BESST_K_EDGE_REDUCE);
                if (runs && runs->packed16)
                    k_edge_gather<true><<<grid, 256, 0, ctx->stream>>>(EA, E, grouped_ptr, ctx->edge_run_ptr.as<u32>(), ctx->run_off.as<u32>(),
                                                                       ctx->run_src.as<u32>(), ctx->run_len.as<u32>(), bv, fishy_sorted, n_fishy,
                                                                       (u32)(2 * ctx->n_large), p.no_score ? 0 : 1);
                else
                    k_edge_gather<false><<<grid, 256, 0, ctx->stream>>>(EA, E, grouped_ptr, ctx->edge_run_ptr.as<u32>(), ctx->run_off.as<u32>(),
                                                                        ctx->run_src.as<u32>(), ctx->run_len.as<u32>(), bv, fishy_sorted, n_fishy,
                                                                        (u32)(2 * ctx->n_large), p.no_score ? 0 : 1);
                BESST_CUDA_TRY(ctx, cudaGetLastError());
            }
            besst_mark(ctx);
            done = true;
        }
    }   // no runs: fall through to the radix bucket

    int n_blocks = (int)((n + HB_TILE - 1) / HB_TILE);
    if (!done && !d_tuples && !runs && n > 0) {   // the links of the last extraction: the radix bucket wants the BAM-ordered array
        rc = besst_ensure_tuples(ctx);
        if (rc) return rc;
        d_tuples = ctx->tuples.as<besst_link_tuple>();
    }
    if (!done) {
    // ---- K3: radix bucket (fallback for input without local order) ----------------------------------
    BESST_CUDA_TRY(ctx, ctx->key_a.ensure(8 * nz));
    BESST_CUDA_TRY(ctx, ctx->key_b.ensure(8 * nz));
    // the BAM-order index rides in the low bits of the sort word when it fits: 8 B per link per pass
    int idx_bits = bits_for((uint64_t)(n > 1 ? n - 1 : 1));
    if (2 * bv + idx_bits > 64) idx_bits = 0;
    int in_b = 0;
    const u32* idx = nullptr;
    if (idx_bits) {
        rc = besst_radix_sort_tuples_packed(ctx, d_tuples, bv, idx_bits, ctx->key_a.as<uint64_t>(), ctx->key_b.as<uint64_t>(), n, &in_b);
        if (rc) return rc;
    } else {
        BESST_CUDA_TRY(ctx, ctx->idx_a.ensure(4 * nz));
        BESST_CUDA_TRY(ctx, ctx->idx_b.ensure(4 * nz));
        rc = besst_radix_sort_tuples(ctx, d_tuples, bv, ctx->key_a.as<uint64_t>(), ctx->key_b.as<uint64_t>(),
                                     ctx->idx_a.as<uint32_t>(), ctx->idx_b.as<uint32_t>(), n, &in_b);
        if (rc) return rc;
        idx = in_b ? ctx->idx_b.as<u32>() : ctx->idx_a.as<u32>();
    }
    const u64* keys = in_b ? ctx->key_b.as<u64>() : ctx->key_a.as<u64>();
    besst_mark(ctx);

    // ---- K4: heads -> row_ptr ------------------------------------------------------------
    BESST_CUDA_TRY(ctx, ctx->block_sums.ensure(4 * (size_t)(n_blocks + 2)));
    u32 n_edges32 = 0;
    if (n > 0) {
        { KTimer kt(ctx, BESST_K_HEADS); k_head_count<<<n_blocks, HB_THREADS, 0, ctx->stream>>>(keys, n, idx_bits, ctx->block_sums.as<u32>()); }
        { KTimer kt(ctx, BESST_K_HEADS); k_scan_blocks<<<1, 1024, 0, ctx->stream>>>(ctx->block_sums.as<u32>(), n_blocks); }
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(&n_edges32, ctx->block_sums.as<u32>() + n_blocks, 4, cudaMemcpyDeviceToHost, ctx->stream));
        BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    rc = alloc_edges((int64_t)n_edges32);
    if (rc) return rc;
    if (E == 0) return finish_empty();
    { KTimer kt(ctx, BESST_K_HEADS); k_head_write<<<n_blocks, HB_THREADS, 0, ctx->stream>>>(keys, n, idx_bits, ctx->block_sums.as<u32>(), EA.row_ptr); }
    besst_mark(ctx);

    {
        long long warps = E;
        int grid = (int)((warps * 32 + 255) / 256);
        if (grid > edge_grid_max) grid = edge_grid_max;
        KTimer kt(ctx, BESST_K_EDGE_REDUCE);
        k_edge_reduce<<<grid, 256, 0, ctx->stream>>>(EA, E, d_tuples, idx, keys, bv, idx_bits, fishy_sorted, n_fishy,
                                                     (u32)(2 * ctx->n_large), p.no_score ? 0 : 1);
        BESST_CUDA_TRY(ctx, cudaGetLastError());
    }
    besst_mark(ctx);
    }
    const size_t Ez = (size_t)(E > 0 ? E : 1);

    // ---- K5/K6: KS + GapEst + score on large-large edges --------------------------------
    if (!p.no_score) {
        // observations are < ins_size_threshold (CreateGraph.py:840), and so is max(obs_v) - obs_v
        double thr = p.ins_size_threshold;
        if (!(thr > 1)) thr = 1;
        if (thr > 2147483647.0) thr = 2147483647.0;
        const int value_bits = bits_for((uint64_t)thr);
        const int edge_bits = bits_for((uint64_t)(E > 1 ? E - 1 : 1));
        const int ls_blocks = (int)((E + LS_TILE - 1) / LS_TILE);
        int64_t n_ll_total = 0;

        // ---- edges with at most KB_G links: in-block sort + evaluation (k_ks_block) ------------------
        // BESST_KS=global sends every edge through the device-wide sorts below (A/B, tests)
        const char* ks_env = getenv("BESST_KS");
        const bool block_ks = !(ks_env && ks_env[0] == 'g') && value_bits + 10 <= 32;   // local edge ids take <= 10 bits
        const int big_thr = block_ks ? KB_G : 0;
        // both link spaces are sized with ONE host round trip: the counts of the small-edge space and of the rest
        BESST_CUDA_TRY(ctx, ctx->ll_off.ensure(4 * (Ez + 2)));
        BESST_CUDA_TRY(ctx, ctx->block_sums.ensure(4 * (size_t)(std::max(ls_blocks, n_blocks) + 2)));
        u64* const hp = reinterpret_cast<u64*>(ctx->host_scalars());
        if (!hp) { ctx->err = "pinned host scratch allocation failed"; return BESST_E_NOMEM; }
        hp[0] = hp[1] = 0;
        if (block_ks) {
            BESST_CUDA_TRY(ctx, ctx->heads.ensure(8 * (size_t)(ls_blocks + 2)));
            u64* bs0 = ctx->heads.as<u64>();
            { KTimer kt(ctx, BESST_K_EDGE_SCORE); k_llc_count<<<ls_blocks, LS_THREADS, 0, ctx->stream>>>(EA.flags, EA.nr, E, KB_G, bs0); }
            { KTimer kt(ctx, BESST_K_EDGE_SCORE); k_scan_blocks64<<<1, 1024, 0, ctx->stream>>>(bs0, ls_blocks); }
            BESST_CUDA_TRY(ctx, cudaMemcpyAsync(&hp[0], bs0 + ls_blocks, 8, cudaMemcpyDeviceToHost, ctx->stream));
        }
        { KTimer kt(ctx, BESST_K_EDGE_SCORE); k_ll_count<<<ls_blocks, LS_THREADS, 0, ctx->stream>>>(EA.flags, EA.nr, E, big_thr, ctx->block_sums.as<u32>()); }
        { KTimer kt(ctx, BESST_K_EDGE_SCORE); k_scan_blocks<<<1, 1024, 0, ctx->stream>>>(ctx->block_sums.as<u32>(), ls_blocks); }
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(&hp[1], ctx->block_sums.as<u32>() + ls_blocks, 4, cudaMemcpyDeviceToHost, ctx->stream));
        BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (block_ks) {
            u64* bs = ctx->heads.as<u64>();
            const u64 tot = hp[0];
            const int64_t n_small = (int64_t)(tot >> 32), n_small_links = (int64_t)(tot & 0xffffffffull);
            n_ll_total += n_small_links;
            if (n_small > 0) {
                BESST_CUDA_TRY(ctx, ctx->ks_key[0].ensure(4 * (size_t)(n_small + 1)));
                BESST_CUDA_TRY(ctx, ctx->ks_key[1].ensure(4 * (size_t)(n_small + 1)));
                u32* ll_edges = ctx->ks_key[0].as<u32>();
                u32* ll_start = ctx->ks_key[1].as<u32>();
                { KTimer kt(ctx, BESST_K_EDGE_SCORE); k_llc_write<<<ls_blocks, LS_THREADS, 0, ctx->stream>>>(EA.flags, EA.nr, E, KB_G, bs, ll_edges, ll_start); }
                KsBlockArgs B;
                B.row_ptr = EA.row_ptr; B.nr = EA.nr; B.sum_u = EA.sum_u; B.obs_sum = EA.obs; B.max_v = EA.max_v;
                B.obs_u = EA.obs_u; B.obs_v = EA.obs_v; B.ll_edges = ll_edges; B.ll_start = ll_start;
                B.n_small = n_small; B.n_links = n_small_links; B.value_bits = value_bits; B.ks = EA.ks;
                const unsigned windows = (unsigned)((n_small_links + KB_W - 1) / KB_W);
                { KTimer kt(ctx, BESST_K_KS_BLOCK); k_ks_block<<<windows, KB_THREADS, 0, ctx->stream>>>(B); }
                BESST_CUDA_TRY(ctx, cudaGetLastError());
            }
        }

        // ---- the rest (edges with more links; every LL edge when the block path is off): LL link space,
        // two device-wide key sorts, co-ranking evaluation ------------------------------------------------
        u32* ll_off = ctx->ll_off.as<u32>();
        const int64_t n_ll = (int64_t)(hp[1] & 0xffffffffull);
        n_ll_total += n_ll;
        ctx->n_ll_links = n_ll_total;
        if (n_ll > 0) {
            { KTimer kt(ctx, BESST_K_EDGE_SCORE); k_ll_write<<<ls_blocks, LS_THREADS, 0, ctx->stream>>>(EA.flags, EA.nr, E, big_thr, ctx->block_sums.as<u32>(), ll_off); }
            const int key_bits = value_bits + edge_bits;
            KsArgs K;
            K.row_ptr = EA.row_ptr; K.ll_off = ll_off; K.flags = EA.flags; K.sum_u = EA.sum_u; K.obs_sum = EA.obs;
            K.max_v = EA.max_v; K.n_edges = E; K.n_ll = n_ll; K.value_bits = value_bits;
            const bool narrow = key_bits <= 32;
            const size_t kb = narrow ? 4 : 8;
            // ks_key[0..1] may hold the compact small-edge lists still being read by k_ks_block: stream order protects them
            BESST_CUDA_TRY(ctx, ctx->ks_key[2].ensure(kb * (size_t)n_ll)); BESST_CUDA_TRY(ctx, ctx->ks_key[3].ensure(kb * (size_t)n_ll));
            BESST_CUDA_TRY(ctx, ctx->ks_key[4].ensure(kb * (size_t)n_ll)); BESST_CUDA_TRY(ctx, ctx->ks_key[5].ensure(kb * (size_t)n_ll));
            long long kgrid = (E * 32 + 255) / 256;
            if (kgrid > (long long)ctx->sm_count * 32) kgrid = (long long)ctx->sm_count * 32;
            const long long chunks = (n_ll + KS_CHUNK - 1) / KS_CHUNK;
            const unsigned egrid = (unsigned)((chunks + 255) / 256);
            int b1 = 0, b2 = 0;
            ctx->sweep_kernel_id = BESST_K_KS_SORT;
            if (narrow) {
                u32 *k1 = ctx->ks_key[2].as<u32>(), *k1t = ctx->ks_key[3].as<u32>(), *k2 = ctx->ks_key[4].as<u32>(), *k2t = ctx->ks_key[5].as<u32>();
                { KTimer kt(ctx, BESST_K_EDGE_SCORE); k_score_keys<u32><<<(unsigned)kgrid, 256, 0, ctx->stream>>>(K, EA.obs_u, EA.obs_v, k1, k2); }
                rc = besst_radix_sort_keys32(ctx, k1, k1t, n_ll, key_bits, &b1); if (rc) return rc;
                rc = besst_radix_sort_keys32(ctx, k2, k2t, n_ll, key_bits, &b2); if (rc) return rc;
                const u32 *s1 = b1 ? k1t : k1, *s2 = b2 ? k2t : k2;
                { KTimer kt(ctx, BESST_K_KS_EVAL); k_ks_eval<u32, 0><<<egrid, 256, 0, ctx->stream>>>(K, s1, s2, EA.ks); }
                { KTimer kt(ctx, BESST_K_KS_EVAL); k_ks_eval<u32, 1><<<egrid, 256, 0, ctx->stream>>>(K, s2, s1, EA.ks); }
            } else {
                u64 *k1 = ctx->ks_key[2].as<u64>(), *k1t = ctx->ks_key[3].as<u64>(), *k2 = ctx->ks_key[4].as<u64>(), *k2t = ctx->ks_key[5].as<u64>();
                { KTimer kt(ctx, BESST_K_EDGE_SCORE); k_score_keys<u64><<<(unsigned)kgrid, 256, 0, ctx->stream>>>(K, EA.obs_u, EA.obs_v, k1, k2); }
                rc = besst_radix_sort_keys(ctx, reinterpret_cast<uint64_t*>(k1), reinterpret_cast<uint64_t*>(k1t), n_ll, key_bits, &b1); if (rc) return rc;
                rc = besst_radix_sort_keys(ctx, reinterpret_cast<uint64_t*>(k2), reinterpret_cast<uint64_t*>(k2t), n_ll, key_bits, &b2); if (rc) return rc;
                const u64 *s1 = b1 ? k1t : k1, *s2 = b2 ? k2t : k2;
                { KTimer kt(ctx, BESST_K_KS_EVAL); k_ks_eval<u64, 0><<<egrid, 256, 0, ctx->stream>>>(K, s1, s2, EA.ks); }
                { KTimer kt(ctx, BESST_K_KS_EVAL); k_ks_eval<u64, 1><<<egrid, 256, 0, ctx->stream>>>(K, s2, s1, EA.ks); }
            }
            ctx->sweep_kernel_id = BESST_K_RADIX_SWEEP;
            BESST_CUDA_TRY(ctx, cudaGetLastError());
        }
        ScoreArgs A;
        A.E = EA; A.n_edges = E; A.scaf_len = ctx->scaf_len.as<int>(); A.c = besst_score_consts(p);
        const unsigned fgrid = (unsigned)((E * 4 + 127) / 128);
        { KTimer kt(ctx, BESST_K_GAPEST); k_edge_finalize<<<fgrid, 128, 0, ctx->stream>>>(A); }
        BESST_CUDA_TRY(ctx, cudaGetLastError());
    }
    besst_mark(ctx);
    ctx->have_graph = true;
    return BESST_OK;
}

int besst_launch_graph(besst_ctx* ctx, const besst_lib_params& p, const besst_link_tuple* d_tuples, int64_t n,
                       const uint64_t* d_fishy, int64_t n_fishy) {
    return launch_graph_impl(ctx, p, d_tuples, n, d_fishy, n_fishy, nullptr);
}

int besst_launch_graph_from_runs(besst_ctx* ctx, const besst_lib_params& p, int64_t n_links, const BesstRunInput& runs,
                                 const uint64_t* d_fishy, int64_t n_fishy) {
    return launch_graph_impl(ctx, p, nullptr, n_links, d_fishy, n_fishy, &runs);
}
