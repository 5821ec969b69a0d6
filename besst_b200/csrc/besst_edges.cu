// K4-K6  sorted link tuples -> CSR edge list with link statistics, KS span
// score and ML gap estimate.  Compiled with -fmad=false so that the fp64
// arithmetic rounds exactly like the reference's Python (and the C oracle).
//
//  K4  segment heads -> row_ptr; one warp per edge reduces its segment:
//      nr_links, obs = sum(o1+o2), obs_sq = sum((o1+o2)^2), first-appearance
//      index, per-scaffold observation lists in BAM order
//      (CreateEdge, CreateGraph.py:842-862), fishy count (:141-163).
//  K5  per large-large edge: sort the two observation lists in shared memory
//      (bitonic), two-sample KS statistic = scipy.stats.ks_2samp(...).statistic
//      as used at CreateGraph.py:582-606.
//  K6  GapEstimator bisection + tr_sk_std_dev (mathstats param_est, call sites
//      CreateGraph.py:537,555) with the four erf/exp arguments of g(d) spread
//      over the 4 lanes of a quad and combined with __shfl_sync; then the score
//      (CreateGraph.py:603-614).  No tensor cores: there is no contraction here.
#include <math.h>

#include "besst_internal.cuh"

int besst_radix_sort_tuples(besst_ctx* ctx, const besst_link_tuple* tuples, int bv, uint64_t* keys_a, uint64_t* keys_b,
                            uint32_t* val_a, uint32_t* val_b, int64_t n, int* result_in_b);

namespace {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int HB_THREADS = 256;
constexpr int HB_ITEMS = 8;
constexpr int HB_TILE = HB_THREADS * HB_ITEMS;

// ---- segment heads ---------------------------------------------------------------
__global__ void __launch_bounds__(HB_THREADS) k_head_count(const u64* __restrict__ keys, long long n, u32* block_sums) {
    __shared__ u32 s_w[HB_THREADS / 32];
    const long long base = (long long)blockIdx.x * HB_TILE;
    u32 c = 0;
#pragma unroll
    for (int i = 0; i < HB_ITEMS; ++i) {
        const long long j = base + i * HB_THREADS + threadIdx.x;
        if (j < n) c += (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
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
    k_head_write(const u64* __restrict__ keys, long long n, const u32* __restrict__ block_sums, long long* row_ptr) {
    __shared__ u32 s_w[HB_THREADS / 32];
    const long long base = (long long)blockIdx.x * HB_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // blocked arrangement so that heads keep their order
    bool head[HB_ITEMS];
    u32 c = 0;
#pragma unroll
    for (int i = 0; i < HB_ITEMS; ++i) {
        const long long j = base + (long long)threadIdx.x * HB_ITEMS + i;
        head[i] = j < n && (j == 0 || keys[j] != keys[j - 1]);
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
};

__global__ void __launch_bounds__(256)
    k_edge_reduce(EdgeArrays E, long long n_edges, const besst_link_tuple* __restrict__ tuples,
                  const u32* __restrict__ sorted_idx, const u64* __restrict__ sorted_keys, int bv,
                  const u64* __restrict__ fishy_sorted, long long n_fishy, u32 n_large2) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long e = warp_global; e < n_edges; e += n_warps) {
        const long long b = E.row_ptr[e], t = E.row_ptr[e + 1];
        long long s = 0, sq = 0;
        for (long long j = b + lane; j < t; j += 32) {
            const u32 idx = __ldg(sorted_idx + j);
            const int4 tp = __ldg(reinterpret_cast<const int4*>(tuples + idx));
            E.obs_u[j] = tp.z;
            E.obs_v[j] = tp.w;
            const long long o = (long long)tp.z + (long long)tp.w;
            s += o;
            sq += o * o;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, off);
            sq += __shfl_xor_sync(0xffffffffu, sq, off);
        }
        if (lane == 0) {
            const u64 key = sorted_keys[b];
            const u32 u = (u32)(key >> bv), v = (u32)(key & ((1ull << bv) - 1ull));
            E.u[e] = u;
            E.v[e] = v;
            E.nr[e] = (int)(t - b);
            E.obs[e] = s;
            E.obs_sq[e] = sq;
            E.first[e] = (long long)sorted_idx[b];
            long long f = 0;
            if (n_fishy > 0) {
                const long long lo = lower_bound_u64(fishy_sorted, n_fishy, key);
                const long long hi = lower_bound_u64(fishy_sorted, n_fishy, key + 1);
                f = hi - lo;
            }
            E.fishy[e] = (int)f;
            E.flags[e] = (u < n_large2 && v < n_large2) ? BESST_EDGE_LL : 0;
            E.gap[e] = 0;
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            E.score[e] = nan; E.ks[e] = nan; E.sd_obs[e] = nan; E.sd_model[e] = nan;
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

// ---- group helpers: G = 32 (warp per edge) or G = blockDim (CTA per edge) ---------------
template <int G>
__device__ __forceinline__ void group_sync() {
    if (G == 32) __syncwarp(); else __syncthreads();
}

template <int G>
__device__ __forceinline__ void bitonic_sort(int* a, int npad, int t) {
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < npad; i += G) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int x = a[i], y = a[ixj];
                    const bool asc = (i & k) == 0;
                    if ((x > y) == asc) { a[i] = y; a[ixj] = x; }
                }
            }
            group_sync<G>();
        }
}

__device__ __forceinline__ int upper_bound_int(const int* a, int n, int key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// number of elements with ((double)a[i] - m) <= z  (a ascending)
__device__ __forceinline__ int count_le_shifted(const int* a, int n, double m, double z) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((double)a[mid] - m <= z) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct ScoreArgs {
    EdgeArrays E;
    long long n_edges;
    const int* scaf_len;
    ScoreConsts c;
    int scoring;
    // big-edge spill list
    int* big_list;        // [0] = count, then edge ids
    long long* big_off;   // scratch offset per big edge
    u64* scratch_used;
    int* scratch;
    long long scratch_cap;
};

constexpr int SC_WARPS = 4;
constexpr int SC_NS = 1024;  // longest list sorted by one warp in shared memory

// KS statistic + score of one edge whose two lists are already loaded into
// sa (obs on edge_u's scaffold) and sb (max - obs on edge_v's scaffold), padded
// to npad with INT_MAX.  Called by the whole group; the result is written by
// thread 0.  red = group reduction scratch (only used when G > 32).
template <int G>
__device__ __forceinline__ void score_edge(const ScoreArgs& A, long long e, int* sa, int* sb, int n, int npad,
                                           long long sum_u, long long sum_y, int t, double* red) {
    bitonic_sort<G>(sa, npad, t);
    bitonic_sort<G>(sb, npad, t);
    const double m1 = (double)sum_u / (double)n;   // l1_mean (:584)
    const double m2 = (double)sum_y / (double)n;   // l2_mean (:591)
    double dmax = 0.0;
    for (int i = t; i < n; i += G) {
        {
            const double z = (double)sa[i] - m1;
            const int k1 = upper_bound_int(sa, n, sa[i]);
            const int k2 = count_le_shifted(sb, n, m2, z);
            const double diff = fabs((double)k1 / (double)n - (double)k2 / (double)n);
            if (diff > dmax) dmax = diff;
        }
        {
            const double z = (double)sb[i] - m2;
            const int k2 = upper_bound_int(sb, n, sb[i]);
            const int k1 = count_le_shifted(sa, n, m1, z);
            const double diff = fabs((double)k1 / (double)n - (double)k2 / (double)n);
            if (diff > dmax) dmax = diff;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, dmax, off);
        if (o > dmax) dmax = o;
    }
    if (G > 32) {
        if ((t & 31) == 0) red[t >> 5] = dmax;
        __syncthreads();
        if (t < 32) {
            dmax = t < G / 32 ? red[t] : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double o = __shfl_xor_sync(0xffffffffu, dmax, off);
                if (o > dmax) dmax = o;
            }
        }
    }
    if (t >= 32) return;  // warp 0 of the group finishes (quad shuffles need a full warp)

    const ScoreConsts& c = A.c;
    const double len1 = (double)A.scaf_len[A.E.u[e] >> 1], len2 = (double)A.scaf_len[A.E.v[e] >> 1];
    const long long obs = A.E.obs[e], obs_sq = A.E.obs_sq[e];
    unsigned char flags = A.E.flags[e] | BESST_EDGE_SCORED;
    const double mean_ = (double)obs / (double)n;                                     // :505
    const double data_observation = ((double)n * c.mean - (double)obs) / (double)n;   // :511
    const bool big = (2 * c.sd < len1) && (2 * c.sd < len2);                          // :536
    const int gap_ml = gap_estimator_quad(c, mean_, len1, len2, big);
    double gap = data_observation;
    if (big) { gap = (double)gap_ml; flags |= BESST_EDGE_BIG; }
    const int gap_int = (int)gap;                                                     // :541
    const bool neg = (-gap > len1) || (-gap > len2);                                  // :542
    const double sd_ml = tr_sk_std_dev_quad(c, len1, len2, gap);
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double score = 0.0, ks_out = nan, sd_obs_out = nan, sd_model_out = nan;
    if (neg) {
        flags |= BESST_EDGE_NEGGAP;
    } else {
        const double std_dev_d_eq_0 = big ? sd_ml : 4294967296.0;                    // :548-558
        double std_dev;
        if (n - 1 == 0) std_dev = 4294967296.0;                                       // :563-564
        else {
            const double q = ((double)obs_sq - (double)n * (mean_ * mean_)) / (double)(n - 1);
            if (q < 0) { std_dev = nan; flags |= BESST_EDGE_CPLX; }
            else std_dev = sqrt(q);                                                   // :561
        }
        const double span_score = n < 5 ? 0.0 : 1 - dmax;                             // :603-606
        double std_dev_score;
        if (std_dev_d_eq_0 == 0.0 || std_dev == 0.0 || std_dev != std_dev) std_dev_score = 0.0;
        else {
            const double x = std_dev / std_dev_d_eq_0, y = std_dev_d_eq_0 / std_dev;
            std_dev_score = y < x ? y : x;
        }
        score = (std_dev_score > 0.5 && span_score > 0.5) ? std_dev_score + span_score : 0.0;  // :614
        ks_out = dmax; sd_obs_out = std_dev; sd_model_out = std_dev_d_eq_0;
    }
    if (t == 0) {
        A.E.gap[e] = gap_int;
        A.E.score[e] = score;
        A.E.ks[e] = ks_out;
        A.E.sd_obs[e] = sd_obs_out;
        A.E.sd_model[e] = sd_model_out;
        A.E.flags[e] = flags;
    }
}

__global__ void __launch_bounds__(SC_WARPS * 32) k_edge_score(const ScoreArgs A) {
    __shared__ int s_a[SC_WARPS][SC_NS];
    __shared__ int s_b[SC_WARPS][SC_NS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long warp_global = (long long)blockIdx.x * SC_WARPS + warp;
    const long long n_warps = (long long)gridDim.x * SC_WARPS;
    for (long long e = warp_global; e < A.n_edges; e += n_warps) {
        if (!(A.E.flags[e] & BESST_EDGE_LL)) continue;
        const int n = A.E.nr[e];
        if (n > SC_NS) {
            if (lane == 0) {
                int npad = 1;
                while (npad < n) npad <<= 1;
                const int slot = atomicAdd(&A.big_list[0], 1);
                A.big_list[1 + slot] = (int)e;
                A.big_off[slot] = (long long)atomicAdd(A.scratch_used, (u64)(2ll * npad));
            }
            continue;
        }
        const long long b = A.E.row_ptr[e];
        int npad = 1;
        while (npad < n) npad <<= 1;
        long long sum_u = 0;
        int max_v = -2147483647 - 1;
        for (int i = lane; i < npad; i += 32) {
            int x = 2147483647, y = -2147483647 - 1;
            if (i < n) { x = A.E.obs_u[b + i]; y = A.E.obs_v[b + i]; sum_u += x; }
            s_a[warp][i] = x;
            s_b[warp][i] = y;
            if (y > max_v) max_v = y;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sum_u += __shfl_xor_sync(0xffffffffu, sum_u, off);
            const int o = __shfl_xor_sync(0xffffffffu, max_v, off);
            if (o > max_v) max_v = o;
        }
        __syncwarp();
        long long sum_y = 0;
        for (int i = lane; i < npad; i += 32) {
            if (i < n) {
                const int y = max_v - s_b[warp][i];   // abs(x - max_obs2), :588-590
                s_b[warp][i] = y;
                sum_y += y;
            } else {
                s_b[warp][i] = 2147483647;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum_y += __shfl_xor_sync(0xffffffffu, sum_y, off);
        __syncwarp();
        score_edge<32>(A, e, s_a[warp], s_b[warp], n, npad, sum_u, sum_y, lane, nullptr);
        __syncwarp();
    }
}

constexpr int SB_THREADS = 256;

// edges with more than SC_NS links: one CTA per edge, lists sorted in global scratch
__global__ void __launch_bounds__(SB_THREADS) k_edge_score_big(const ScoreArgs A) {
    __shared__ double s_red[SB_THREADS / 32];
    __shared__ long long s_sum[SB_THREADS / 32];
    __shared__ int s_max[SB_THREADS / 32];
    const int n_big = A.big_list[0];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int k = blockIdx.x; k < n_big; k += gridDim.x) {
        const long long e = A.big_list[1 + k];
        const int n = A.E.nr[e];
        int npad = 1;
        while (npad < n) npad <<= 1;
        const long long off = A.big_off[k];
        if (off + 2ll * npad > A.scratch_cap) continue;  // cannot happen: cap >= 4*n_links
        int* sa = A.scratch + off;
        int* sb = sa + npad;
        const long long b = A.E.row_ptr[e];
        long long sum_u = 0;
        int max_v = -2147483647 - 1;
        for (int i = t; i < npad; i += SB_THREADS) {
            int x = 2147483647, y = -2147483647 - 1;
            if (i < n) { x = A.E.obs_u[b + i]; y = A.E.obs_v[b + i]; sum_u += x; }
            sa[i] = x;
            sb[i] = y;
            if (y > max_v) max_v = y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum_u += __shfl_xor_sync(0xffffffffu, sum_u, o);
            const int m = __shfl_xor_sync(0xffffffffu, max_v, o);
            if (m > max_v) max_v = m;
        }
        if (lane == 0) { s_sum[warp] = sum_u; s_max[warp] = max_v; }
        __syncthreads();
        sum_u = 0;
        for (int w = 0; w < SB_THREADS / 32; ++w) { sum_u += s_sum[w]; if (s_max[w] > max_v) max_v = s_max[w]; }
        __syncthreads();
        long long sum_y = 0;
        for (int i = t; i < npad; i += SB_THREADS) {
            if (i < n) { const int y = max_v - sb[i]; sb[i] = y; sum_y += y; }
            else sb[i] = 2147483647;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum_y += __shfl_xor_sync(0xffffffffu, sum_y, o);
        if (lane == 0) s_sum[warp] = sum_y;
        __syncthreads();
        sum_y = 0;
        for (int w = 0; w < SB_THREADS / 32; ++w) sum_y += s_sum[w];
        __syncthreads();
        score_edge<SB_THREADS>(A, e, sa, sb, n, npad, sum_u, sum_y, t, s_red);
        __syncthreads();
    }
}

// ---- batched GapEstimator + tr_sk_std_dev: one quad per item -------------------------------
__global__ void __launch_bounds__(128)
    k_gapest_batch(const ScoreConsts c, const double* __restrict__ mean_obs, const int* __restrict__ len1,
                   const int* __restrict__ len2, long long n, int* gap_out, double* sd_out) {
    const long long quad = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool active = quad < n;
    const double mo = active ? mean_obs[quad] : 0.0;
    const double l1 = active ? (double)len1[quad] : 1.0, l2 = active ? (double)len2[quad] : 1.0;
    const int gap = gap_estimator_quad(c, mo, l1, l2, active);
    const double sd = tr_sk_std_dev_quad(c, l1, l2, (double)gap);
    if (active && (threadIdx.x & 3) == 0) {
        gap_out[quad] = gap;
        if (sd_out) sd_out[quad] = sd;
    }
}

__global__ void __launch_bounds__(128)
    k_trsk_sd_batch(const ScoreConsts c, const double* __restrict__ gap, const int* __restrict__ len1,
                    const int* __restrict__ len2, long long n, double* sd_out) {
    const long long quad = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool active = quad < n;
    const double d = active ? gap[quad] : 0.0;
    const double l1 = active ? (double)len1[quad] : 1.0, l2 = active ? (double)len2[quad] : 1.0;
    const double sd = tr_sk_std_dev_quad(c, l1, l2, d);
    if (active && (threadIdx.x & 3) == 0) sd_out[quad] = sd;
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

int besst_launch_gapest(besst_ctx* ctx, const besst_lib_params& p, const double* d_mean_obs, const int32_t* d_len1,
                        const int32_t* d_len2, int64_t n, int32_t* d_gap, double* d_sd) {
    if (n == 0) return BESST_OK;
    const ScoreConsts c = besst_score_consts(p);
    const long long threads = n * 4;
    const int grid = (int)((threads + 127) / 128);
    { KTimer kt(ctx, BESST_K_GAPEST); k_gapest_batch<<<grid, 128, 0, ctx->stream>>>(c, d_mean_obs, d_len1, d_len2, n, d_gap, d_sd); }
    BESST_CUDA_TRY(ctx, cudaGetLastError());
    return BESST_OK;
}

int besst_launch_trsk_sd(besst_ctx* ctx, const besst_lib_params& p, const double* d_gap, const int32_t* d_len1,
                         const int32_t* d_len2, int64_t n, double* d_sd) {
    if (n == 0) return BESST_OK;
    const ScoreConsts c = besst_score_consts(p);
    const long long threads = n * 4;
    const int grid = (int)((threads + 127) / 128);
    { KTimer kt(ctx, BESST_K_GAPEST); k_trsk_sd_batch<<<grid, 128, 0, ctx->stream>>>(c, d_gap, d_len1, d_len2, n, d_sd); }
    BESST_CUDA_TRY(ctx, cudaGetLastError());
    return BESST_OK;
}

int besst_launch_graph(besst_ctx* ctx, const besst_lib_params& p, const besst_link_tuple* d_tuples, int64_t n,
                       const uint64_t* d_fishy, int64_t n_fishy) {
    ctx->have_graph = false;
    ctx->n_links = n;
    ctx->n_edges = 0;
    ctx->last_params = p;
    const int bv = bits_for((uint64_t)(2 * ctx->n_scaffolds > 0 ? 2 * ctx->n_scaffolds - 1 : 1));
    const size_t nz = (size_t)(n > 0 ? n : 1);

    // ---- K3: radix bucket -------------------------------------------------------------
    BESST_CUDA_TRY(ctx, ctx->key_a.ensure(8 * nz));
    BESST_CUDA_TRY(ctx, ctx->key_b.ensure(8 * nz));
    BESST_CUDA_TRY(ctx, ctx->idx_a.ensure(4 * nz));
    BESST_CUDA_TRY(ctx, ctx->idx_b.ensure(4 * nz));
    int in_b = 0;
    int rc = besst_radix_sort_tuples(ctx, d_tuples, bv, ctx->key_a.as<uint64_t>(), ctx->key_b.as<uint64_t>(),
                                     ctx->idx_a.as<uint32_t>(), ctx->idx_b.as<uint32_t>(), n, &in_b);
    if (rc) return rc;
    const u64* keys = in_b ? ctx->key_b.as<u64>() : ctx->key_a.as<u64>();
    const u32* idx = in_b ? ctx->idx_b.as<u32>() : ctx->idx_a.as<u32>();
    besst_mark(ctx);

    // fishy pairs: rekey, sort (keys only)
    const u64* fishy_sorted = nullptr;
    if (n_fishy > 0) {
        BESST_CUDA_TRY(ctx, ctx->fishy_sorted.ensure(8 * (size_t)n_fishy));
        BESST_CUDA_TRY(ctx, ctx->fishy_tmp.ensure(8 * (size_t)n_fishy));
        { KTimer kt(ctx, BESST_K_FISHY); k_fishy_rekey<<<(int)((n_fishy + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const u64*>(d_fishy), ctx->fishy_sorted.as<u64>(), n_fishy, bv); }
        int fb = 0;
        rc = besst_radix_sort_keys(ctx, ctx->fishy_sorted.as<uint64_t>(), ctx->fishy_tmp.as<uint64_t>(), n_fishy, 2 * bv, &fb);
        if (rc) return rc;
        fishy_sorted = fb ? ctx->fishy_tmp.as<u64>() : ctx->fishy_sorted.as<u64>();
    }

    // ---- K4: heads -> row_ptr ------------------------------------------------------------
    const int n_blocks = (int)((n + HB_TILE - 1) / HB_TILE);
    BESST_CUDA_TRY(ctx, ctx->block_sums.ensure(4 * (size_t)(n_blocks + 2)));
    u32 n_edges32 = 0;
    if (n > 0) {
        { KTimer kt(ctx, BESST_K_HEADS); k_head_count<<<n_blocks, HB_THREADS, 0, ctx->stream>>>(keys, n, ctx->block_sums.as<u32>()); }
        { KTimer kt(ctx, BESST_K_HEADS); k_scan_blocks<<<1, 1024, 0, ctx->stream>>>(ctx->block_sums.as<u32>(), n_blocks); }
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(&n_edges32, ctx->block_sums.as<u32>() + n_blocks, 4, cudaMemcpyDeviceToHost, ctx->stream));
        BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    const int64_t E = n_edges32;
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
    EdgeArrays EA;
    EA.u = ctx->e_u.as<u32>(); EA.v = ctx->e_v.as<u32>(); EA.nr = ctx->e_nr.as<int>();
    EA.obs = ctx->e_obs.as<long long>(); EA.obs_sq = ctx->e_obs_sq.as<long long>(); EA.first = ctx->e_first.as<long long>();
    EA.row_ptr = ctx->e_row_ptr.as<long long>(); EA.gap = ctx->e_gap.as<int>(); EA.score = ctx->e_score.as<double>();
    EA.ks = ctx->e_ks.as<double>(); EA.sd_obs = ctx->e_sd_obs.as<double>(); EA.sd_model = ctx->e_sd_model.as<double>();
    EA.fishy = ctx->e_fishy.as<int>(); EA.flags = ctx->e_flags.as<unsigned char>();
    EA.obs_u = ctx->l_obs_u.as<int>(); EA.obs_v = ctx->l_obs_v.as<int>();
    if (E == 0) {
        const long long zero = 0;
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(EA.row_ptr, &zero, 8, cudaMemcpyHostToDevice, ctx->stream));
        BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        besst_mark(ctx); besst_mark(ctx); besst_mark(ctx);
        ctx->have_graph = true;
        return BESST_OK;
    }
    { KTimer kt(ctx, BESST_K_HEADS); k_head_write<<<n_blocks, HB_THREADS, 0, ctx->stream>>>(keys, n, ctx->block_sums.as<u32>(), EA.row_ptr); }
    besst_mark(ctx);

    {
        long long warps = E;
        int grid = (int)((warps * 32 + 255) / 256);
        const int max_grid = ctx->sm_count * 32;
        if (grid > max_grid) grid = max_grid;
        KTimer kt(ctx, BESST_K_EDGE_REDUCE);
        k_edge_reduce<<<grid, 256, 0, ctx->stream>>>(EA, E, d_tuples, idx, keys, bv, fishy_sorted, n_fishy,
                                                     (u32)(2 * ctx->n_large));
        BESST_CUDA_TRY(ctx, cudaGetLastError());
    }
    besst_mark(ctx);

    // ---- K5/K6: KS + GapEst + score on large-large edges --------------------------------
    if (!p.no_score) {
        BESST_CUDA_TRY(ctx, ctx->big_list.ensure(4 * (Ez + 1) + 8 * Ez + 16));
        BESST_CUDA_TRY(ctx, ctx->big_scratch.ensure(4 * (4 * nz + 16)));
        ScoreArgs A;
        A.E = EA; A.n_edges = E; A.scaf_len = ctx->scaf_len.as<int>(); A.c = besst_score_consts(p); A.scoring = 1;
        unsigned char* bl = ctx->big_list.as<unsigned char>();
        A.scratch_used = reinterpret_cast<u64*>(bl);
        A.big_off = reinterpret_cast<long long*>(bl + 8);
        A.big_list = reinterpret_cast<int*>(bl + 8 + 8 * Ez);
        A.scratch = ctx->big_scratch.as<int>();
        A.scratch_cap = (long long)(4 * nz + 16);
        BESST_CUDA_TRY(ctx, cudaMemsetAsync(A.scratch_used, 0, 8, ctx->stream));
        BESST_CUDA_TRY(ctx, cudaMemsetAsync(A.big_list, 0, 4, ctx->stream));
        int grid = (int)((E + SC_WARPS - 1) / SC_WARPS);
        const int max_grid = ctx->sm_count * 16;
        if (grid > max_grid) grid = max_grid;
        { KTimer kt(ctx, BESST_K_EDGE_SCORE); k_edge_score<<<grid, SC_WARPS * 32, 0, ctx->stream>>>(A); }
        { KTimer kt(ctx, BESST_K_EDGE_SCORE); k_edge_score_big<<<ctx->sm_count, SB_THREADS, 0, ctx->stream>>>(A); }
        BESST_CUDA_TRY(ctx, cudaGetLastError());
    }
    besst_mark(ctx);
    ctx->have_graph = true;
    return BESST_OK;
}
