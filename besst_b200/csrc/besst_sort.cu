// K3  radix bucket: stable LSD radix sort of 64-bit edge keys with a 32-bit
// payload (the tuple's BAM-order index), one-sweep style.
//
// This is the GPU replacement for the dict-of-dict upsert of CreateEdge
// (CreateGraph.py:842-862): instead of hashing (scaffold, side) tuples one link
// at a time, all accepted links are bucketed by their canonical edge key
// `(u << bits) | v`; because the sort is stable and the input is in BAM order,
// every edge's links stay in BAM order (the order of the reference's
// `observations` lists) and the first link of a segment is the edge's first
// appearance (the networkx insertion order).
//
// Per pass (8-bit digit) each key is read once and written once:
//  * one up-front kernel builds the digit histograms of ALL passes from a single
//    read of the keys (__match_any_sync warp-aggregated shared-memory atomics);
//  * the pass kernel ranks a 4096-key tile with per-warp match_any multisplit,
//    gets its global digit offsets by decoupled look-back over per-tile digit
//    counts, reorders the tile through shared memory and writes digit runs with
//    coalesced stores.
#include "besst_internal.cuh"

namespace {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_RADIX = 256;
constexpr int RS_MAX_PASSES = 8;
constexpr u32 RS_AGG = 1u << 30, RS_INC = 2u << 30, RS_VAL = (1u << 30) - 1;

__device__ __forceinline__ u32 ld_vol32(const u32* p) { return *reinterpret_cast<const volatile u32*>(p); }
__device__ __forceinline__ void st_vol32(u32* p, u32 v) { *reinterpret_cast<volatile u32*>(p) = v; }

// ---- histograms of all passes in one read --------------------------------------
template <typename KeyT, bool FROM_TUPLES>
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const KeyT* __restrict__ keys,
                                                            const besst_link_tuple* __restrict__ tuples, int bv,
                                                            long long n, int passes, u32* __restrict__ ghist) {
    __shared__ u32 sh[RS_MAX_PASSES * RS_RADIX];
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += RS_THREADS) sh[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * RS_THREADS;
    const long long n_round = (n + 31) / 32 * 32;
    for (long long i = (long long)blockIdx.x * RS_THREADS + threadIdx.x; i < n_round; i += stride) {
        const bool valid = i < n;
        u64 key = 0;
        if (valid) {
            if (FROM_TUPLES) {
                const uint2 uv = __ldg(reinterpret_cast<const uint2*>(tuples + i));
                key = ((u64)uv.x << bv) | uv.y;
            } else {
                key = (u64)__ldg(keys + i);
            }
        }
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            for (int p = 0; p < passes; ++p) {
                const u32 d = (u32)(key >> (8 * p)) & 255u;
                const unsigned peers = __match_any_sync(vmask, d);
                if (lane == __ffs(peers) - 1) atomicAdd(&sh[p * RS_RADIX + d], (u32)__popc(peers));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += RS_THREADS)
        if (sh[i]) atomicAdd(&ghist[i], sh[i]);
}

// exclusive scan of each pass's 256-bin histogram (one CTA per pass)
__global__ void __launch_bounds__(RS_RADIX) k_radix_scan_hist(u32* ghist) {
    __shared__ u32 s[RS_RADIX];
    u32* h = ghist + blockIdx.x * RS_RADIX;
    const u32 v = h[threadIdx.x];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < RS_RADIX; off <<= 1) {
        u32 t = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    h[threadIdx.x] = s[threadIdx.x] - v;
}

template <typename KeyT, bool HAS_VAL>
struct SweepSmem {
    KeyT keys[RS_TILE];
    u32 vals[HAS_VAL ? RS_TILE : 1];
    u32 warp_hist[RS_WARPS][RS_RADIX];
    u32 digit_start[RS_RADIX];
    long long gbase[RS_RADIX];
    u32 warp_sum[RS_WARPS];
    int tile;
};

// FROM_TUPLES: the first pass reads (u, v) straight from the link tuples.  With idx_bits > 0 the
// BAM-order index is packed into the low bits of the key, (((u << bv) | v) << idx_bits) | index, and
// there is no separate payload array: one 8-byte word per link per pass instead of 12.
template <typename KeyT, bool FROM_TUPLES, bool HAS_VAL>
__global__ void __launch_bounds__(RS_THREADS, 3)
    k_radix_sweep(const KeyT* __restrict__ in_keys, const u32* __restrict__ in_vals,
                  const besst_link_tuple* __restrict__ tuples, int bv, int idx_bits, KeyT* __restrict__ out_keys,
                  u32* __restrict__ out_vals, long long n, int shift, const u32* __restrict__ gbase,
                  u32* status, u32* ticket, int n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SweepSmem<KeyT, HAS_VAL>& S = *reinterpret_cast<SweepSmem<KeyT, HAS_VAL>*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) S.tile = (int)atomicAdd(ticket, 1u);
        for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&S.warp_hist[0][0])[i] = 0;
        __syncthreads();
        const int tile = S.tile;
        if (tile >= n_tiles) break;
        const long long tile_base = (long long)tile * RS_TILE;
        const long long rem = n - tile_base;
        const int tile_count = rem < RS_TILE ? (int)rem : RS_TILE;

        KeyT key[RS_ITEMS];
        u32 val[HAS_VAL ? RS_ITEMS : 1];
        unsigned short rank[RS_ITEMS];
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const int off = warp * (32 * RS_ITEMS) + i * 32 + lane;
            key[i] = (KeyT)~0ull;
            if (HAS_VAL) val[i] = 0;
            if (off < tile_count) {
                if (FROM_TUPLES) {
                    const uint2 uv = __ldg(reinterpret_cast<const uint2*>(tuples + tile_base + off));
                    const u64 k = ((u64)uv.x << bv) | uv.y;
                    if (HAS_VAL) { key[i] = (KeyT)k; val[i] = (u32)(tile_base + off); }
                    else key[i] = (KeyT)((k << idx_bits) | (u64)(tile_base + off));
                } else {
                    key[i] = __ldg(in_keys + tile_base + off);
                    if (HAS_VAL) val[i] = __ldg(in_vals + tile_base + off);
                }
            }
        }
        // per-warp stable multisplit
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const int off = warp * (32 * RS_ITEMS) + i * 32 + lane;
            const bool valid = off < tile_count;
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            const u32 d = (u32)(key[i] >> shift) & 255u;
            unsigned peers = 0;
            u32 pre = 0;
            if (valid) {
                peers = __match_any_sync(vmask, d);
                pre = S.warp_hist[warp][d];
            }
            __syncwarp();
            if (valid) {
                rank[i] = (unsigned short)(pre + __popc(peers & lt_mask));
                if (lane == __ffs(peers) - 1) S.warp_hist[warp][d] = pre + (u32)__popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        // digit totals of the tile and per-warp offsets (thread d owns digit d); publish the tile's
        // digit counts as early as possible
        u32 tile_hist, dstart;
        u32* st = status + (size_t)tile * RS_RADIX + threadIdx.x;
        {
            const int d = threadIdx.x;
            u32 run = 0;
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) {
                const u32 t = S.warp_hist[w][d];
                S.warp_hist[w][d] = run;
                run += t;
            }
            tile_hist = run;
            st_vol32(st, (tile == 0 ? RS_INC : RS_AGG) | tile_hist);
            // exclusive scan over digits
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
            for (int w = 0; w < RS_WARPS; ++w)
                if (w < warp) wbase += S.warp_sum[w];
            dstart = wbase + incl - run;
            S.digit_start[d] = dstart;
        }
        __syncthreads();
        // local reorder through shared memory (gives the predecessors time to publish)
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const int off = warp * (32 * RS_ITEMS) + i * 32 + lane;
            if (off < tile_count) {
                const u32 d = (u32)(key[i] >> shift) & 255u;
                const u32 lp = S.digit_start[d] + S.warp_hist[warp][d] + rank[i];
                S.keys[lp] = key[i];
                if (HAS_VAL) S.vals[lp] = val[i];
            }
        }
        // decoupled look-back for digit d = threadIdx.x, four predecessors per round trip
        {
            const int d = threadIdx.x;
            u32 excl = 0;
            int p = tile - 1;
            while (p >= 0) {
                u32 w[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) w[k] = (p - k >= 0) ? ld_vol32(status + (size_t)(p - k) * RS_RADIX + d) : RS_INC;
                int used = 0;
                bool done = false;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (done || used != k) continue;
                    const u32 flag = w[k] & ~RS_VAL;
                    if (flag == 0) continue;          // not published yet: re-poll from here
                    excl += w[k] & RS_VAL;
                    used = k + 1;
                    if (flag == RS_INC) done = true;
                }
                if (done) break;
                p -= used;
            }
            if (tile > 0) st_vol32(st, RS_INC | (excl + tile_hist));
            S.gbase[d] = (long long)gbase[d] + (long long)excl - (long long)dstart;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < tile_count; j += RS_THREADS) {
            const KeyT k = S.keys[j];
            const u32 d = (u32)(k >> shift) & 255u;
            const long long dest = S.gbase[d] + j;
            out_keys[dest] = k;
            if (HAS_VAL) out_vals[dest] = S.vals[j];
        }
    }
}

template <typename KeyT, bool FROM_TUPLES, bool HAS_VAL>
int sort_impl(besst_ctx* ctx, const besst_link_tuple* tuples, int bv, int idx_bits, KeyT* keys_a, KeyT* keys_b, u32* val_a,
              u32* val_b, int64_t n, int key_bits, int* result_in_b) {
    *result_in_b = 0;
    int passes = (key_bits + 7) / 8;
    if (passes < 1) passes = 1;
    if (passes > RS_MAX_PASSES) { ctx->err = "radix sort: key too wide"; return BESST_E_INVALID; }
    if (n >= (1ll << 30)) { ctx->err = "radix sort: more than 2^30 keys in one call"; return BESST_E_INVALID; }
    if (n == 0) return BESST_OK;
    const int n_tiles = (int)((n + RS_TILE - 1) / RS_TILE);
    BESST_CUDA_TRY(ctx, ctx->hist.ensure(sizeof(u32) * (RS_MAX_PASSES * RS_RADIX + 64)));
    BESST_CUDA_TRY(ctx, ctx->sort_state.ensure(sizeof(u32) * (size_t)n_tiles * RS_RADIX));
    u32* ghist = ctx->hist.as<u32>();
    u32* tickets = ghist + RS_MAX_PASSES * RS_RADIX;
    BESST_CUDA_TRY(ctx, cudaMemsetAsync(ghist, 0, sizeof(u32) * (RS_MAX_PASSES * RS_RADIX + 64), ctx->stream));
    int hgrid = ctx->sm_count * 8;
    const long long max_blocks = (n + RS_THREADS - 1) / RS_THREADS;
    if (hgrid > max_blocks) hgrid = (int)max_blocks;
    { KTimer kt(ctx, BESST_K_RADIX_HIST); k_radix_hist<KeyT, FROM_TUPLES><<<hgrid, RS_THREADS, 0, ctx->stream>>>(keys_a, tuples, bv, n, passes, ghist); }
    { KTimer kt(ctx, BESST_K_RADIX_SCAN); k_radix_scan_hist<<<passes, RS_RADIX, 0, ctx->stream>>>(ghist); }
    BESST_CUDA_TRY(ctx, cudaGetLastError());

    const size_t smem = sizeof(SweepSmem<KeyT, HAS_VAL>);
    // every call: cheap, and a process may drive several devices (one ctx each) with the same kernels
    if (FROM_TUPLES) cudaFuncSetAttribute(k_radix_sweep<KeyT, FROM_TUPLES, HAS_VAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_radix_sweep<KeyT, false, HAS_VAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_radix_sweep<KeyT, false, HAS_VAL>, RS_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    int grid = ctx->sm_count * per_sm;
    if (grid > n_tiles) grid = n_tiles;

    const KeyT* in_k = keys_a;
    const u32* in_v = val_a;
    KeyT* out_k = keys_b;
    u32* out_v = val_b;
    for (int p = 0; p < passes; ++p) {
        BESST_CUDA_TRY(ctx, cudaMemsetAsync(ctx->sort_state.p, 0, sizeof(u32) * (size_t)n_tiles * RS_RADIX, ctx->stream));
        {
        KTimer kt(ctx, ctx->sweep_kernel_id);
        if (p == 0 && FROM_TUPLES)
            k_radix_sweep<KeyT, FROM_TUPLES, HAS_VAL><<<grid, RS_THREADS, smem, ctx->stream>>>(
                nullptr, nullptr, tuples, bv, idx_bits, out_k, out_v, n, idx_bits, ghist, ctx->sort_state.as<u32>(), tickets + p, n_tiles);
        else
            k_radix_sweep<KeyT, false, HAS_VAL><<<grid, RS_THREADS, smem, ctx->stream>>>(
                in_k, in_v, nullptr, bv, idx_bits, out_k, out_v, n, idx_bits + 8 * p, ghist + p * RS_RADIX, ctx->sort_state.as<u32>(),
                tickets + p, n_tiles);
        }
        BESST_CUDA_TRY(ctx, cudaGetLastError());
        // ping-pong: after the first pass from tuples the data lives in (keys_b, val_b)
        const KeyT* nk = out_k;
        const u32* nv = out_v;
        out_k = (out_k == keys_b) ? keys_a : keys_b;
        out_v = (out_v == val_b) ? val_a : val_b;
        in_k = nk;
        in_v = nv;
    }
    *result_in_b = (in_k == keys_b) ? 1 : 0;
    return BESST_OK;
}

}  // namespace

int besst_radix_sort_pairs(besst_ctx* ctx, uint64_t* keys_a, uint64_t* keys_b, uint32_t* val_a, uint32_t* val_b,
                           int64_t n, int key_bits, int* result_in_b) {
    return sort_impl<u64, false, true>(ctx, nullptr, 0, 0, reinterpret_cast<u64*>(keys_a), reinterpret_cast<u64*>(keys_b), val_a,
                                       val_b, n, key_bits, result_in_b);
}

int besst_radix_sort_keys(besst_ctx* ctx, uint64_t* keys_a, uint64_t* keys_b, int64_t n, int key_bits,
                          int* result_in_b) {
    return sort_impl<u64, false, false>(ctx, nullptr, 0, 0, reinterpret_cast<u64*>(keys_a), reinterpret_cast<u64*>(keys_b),
                                        nullptr, nullptr, n, key_bits, result_in_b);
}

int besst_radix_sort_keys32(besst_ctx* ctx, uint32_t* keys_a, uint32_t* keys_b, int64_t n, int key_bits, int* result_in_b) {
    return sort_impl<u32, false, false>(ctx, nullptr, 0, 0, keys_a, keys_b, nullptr, nullptr, n, key_bits, result_in_b);
}

// tuples (BAM order) -> sorted (key, original index); key = (u << bv) | v
int besst_radix_sort_tuples(besst_ctx* ctx, const besst_link_tuple* tuples, int bv, uint64_t* keys_a, uint64_t* keys_b,
                            uint32_t* val_a, uint32_t* val_b, int64_t n, int* result_in_b) {
    return sort_impl<u64, true, true>(ctx, tuples, bv, 0, reinterpret_cast<u64*>(keys_a), reinterpret_cast<u64*>(keys_b), val_a,
                                      val_b, n, 2 * bv, result_in_b);
}

// tuples (BAM order) -> sorted packed words (((u << bv) | v) << idx_bits) | original index; needs
// 2 * bv + idx_bits <= 64 and n <= 2^idx_bits
int besst_radix_sort_tuples_packed(besst_ctx* ctx, const besst_link_tuple* tuples, int bv, int idx_bits, uint64_t* keys_a,
                                   uint64_t* keys_b, int64_t n, int* result_in_b) {
    return sort_impl<u64, true, false>(ctx, tuples, bv, idx_bits, reinterpret_cast<u64*>(keys_a), reinterpret_cast<u64*>(keys_b),
                                       nullptr, nullptr, n, 2 * bv, result_in_b);
}
