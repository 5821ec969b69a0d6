// K7  library metrics: the capped, BAM-ordered sampling passes of
// libmetrics.get_metrics (libmetrics.py:283-304) and get_contamination_metrics
// (:49-84) as one streaming pair of kernels, plus the O(bins) statistics that
// follow them (:316-343 trim loop, getdistr :141-223, :88-110) on the host side
// of the shim.
//
// The reference appends |tlen| of qualifying read2 records to a Python list
// until it holds 1e6 samples, and separately walks the first 1e6 records that
// sit on the 1000 longest contigs.  Both cuts are prefix cuts in BAM order, so
// the GPU version ranks qualifying records with a block scan + a scan of the
// per-tile counts and lets every record with rank < 1e6 add 1 to an integer
// histogram over |tlen| (order-free, exact).  Everything downstream only needs
// the multiset of samples, i.e. the histogram: mean, sd, the
// AdjustInsertsizeDist trim loop, skewness and GetDistr's weighting are
// evaluated over bins (fp64), which differs from the reference's sample-order
// summation by rounding only (tests: <= 1e-9 relative).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "besst_internal.cuh"

namespace {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int MT_THREADS = 256;
constexpr int MT_ITEMS = 8;
constexpr int MT_TILE = MT_THREADS * MT_ITEMS;
constexpr long long MT_CAP = 1000000;
constexpr long long MT_CHUNK = 1ll << 26;

struct MParams {
    DeviceRecords rec;
    const int4* rows;  // two int4 per contig row; .y of the second = in_largest... see row layout
    int n_contigs;
    int orientation, min_mapq, want_isize;
    double read_len;
    long long lo, hi;  // record range of this chunk
};

// bam_parser.is_proper_aligned_unique_innie / _outie (bam_parser.py:22-29)
__device__ __forceinline__ bool proper_pair(unsigned flag, int tid, int mtid, int tlen, int mapq, int thr, bool innie) {
    const bool rev = flag & 0x10u, mrev = flag & 0x20u, read2 = flag & 0x80u;
    const bool mate_unmapped = flag & 0x8u, secondary = flag & 0x100u;
    const bool neg = innie ? tlen < 0 : tlen > 0, pos = innie ? tlen > 0 : tlen < 0;
    const bool geom = read2 && tid == mtid && ((rev && !mrev && neg) || (!rev && mrev && pos));
    return geom && !mate_unmapped && mapq > thr && !secondary;
}

struct Cls {
    bool isize, scope, mapped, cont;
    int v;
};

__device__ __forceinline__ Cls classify(const MParams& P, long long j) {
    Cls c = {false, false, false, false, 0};
    if (j >= P.hi) return c;
    const int tid = __ldg(P.rec.tid + j);
    if (tid < 0 || tid >= P.n_contigs) return c;
    const int4 rb = __ldg(P.rows + 2 * (long long)tid + 1);  // length, scaf_length, in_largest, reserved
    if (!rb.z) return c;
    const unsigned flag = __ldg(P.rec.flag + j);
    const int mtid = __ldg(P.rec.mtid + j), tlen = __ldg(P.rec.tlen + j), mapq = __ldg(P.rec.mapq + j);
    const bool fr = P.orientation == BESST_ORIENT_FR;
    c.scope = true;                                   // libmetrics.py:65
    c.mapped = !(flag & 0x4u);                        // :67-68
    c.v = tlen < 0 ? -tlen : tlen;
    if (P.want_isize) c.isize = proper_pair(flag, tid, mtid, tlen, mapq, P.min_mapq, fr);       // :294-301
    if (proper_pair(flag, tid, mtid, tlen, mapq, P.min_mapq, !fr)) {                           // :71-81
        const double frag = fr ? (double)c.v + 2 * P.read_len : (double)c.v;
        c.cont = P.read_len < frag;
    }
    return c;
}

__global__ void __launch_bounds__(MT_THREADS) k_metrics_count(const MParams P, u32* cnt_isize, u32* cnt_scope) {
    __shared__ u32 s_a[MT_THREADS / 32], s_b[MT_THREADS / 32];
    const long long base = P.lo + (long long)blockIdx.x * MT_TILE;
    u32 a = 0, b = 0;
#pragma unroll
    for (int i = 0; i < MT_ITEMS; ++i) {
        const Cls c = classify(P, base + i * MT_THREADS + threadIdx.x);
        a += c.isize;
        b += c.scope;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 ta = 0, tb = 0;
        for (int w = 0; w < MT_THREADS / 32; ++w) { ta += s_a[w]; tb += s_b[w]; }
        cnt_isize[blockIdx.x] = ta;
        cnt_scope[blockIdx.x] = tb;
    }
}

// single-CTA exclusive scan of both per-tile count arrays; totals at [n_tiles]
__global__ void __launch_bounds__(1024) k_metrics_scan(u32* a, u32* b, int n_tiles) {
    __shared__ u32 s_w[2][32];
    __shared__ u32 s_carry[2];
    if (threadIdx.x < 2) s_carry[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        u32 v[2] = {i < n_tiles ? a[i] : 0, i < n_tiles ? b[i] : 0};
        u32 incl[2] = {v[0], v[1]};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const u32 t = __shfl_up_sync(0xffffffffu, incl[k], off);
                if (lane >= off) incl[k] += t;
            }
            if (lane == 31) s_w[k][warp] = incl[k];
        }
        __syncthreads();
        if (warp < 2) {
            const u32 w = s_w[warp][lane];
            u32 wi = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const u32 t = __shfl_up_sync(0xffffffffu, wi, off);
                if (lane >= off) wi += t;
            }
            s_w[warp][lane] = wi - w;
        }
        __syncthreads();
        const u32 c0 = s_carry[0], c1 = s_carry[1];
        const u32 e0 = c0 + s_w[0][warp] + incl[0] - v[0], e1 = c1 + s_w[1][warp] + incl[1] - v[1];
        if (i < n_tiles) { a[i] = e0; b[i] = e1; }
        __syncthreads();
        if (threadIdx.x == 1023) { s_carry[0] = e0 + v[0]; s_carry[1] = e1 + v[1]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { a[n_tiles] = s_carry[0]; b[n_tiles] = s_carry[1]; }
}

// out[0] = counter_total (mapped records in scope), out[1] = 1 + index of the
// record holding the last isize sample, out[2] = same for the scope cut,
// out[3] = largest |tlen| among the counted samples (a value >= n_bins means
// the histogram was too short: the host re-runs with more bins)
__global__ void __launch_bounds__(MT_THREADS)
    k_metrics_hist(const MParams P, const u32* __restrict__ pre_isize, const u32* __restrict__ pre_scope, long long base_isize,
                   long long base_scope, u32* hist_isize, u32* hist_cont, int n_bins, u64* out) {
    __shared__ u32 s_a[MT_THREADS / 32], s_b[MT_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long first = P.lo + (long long)blockIdx.x * MT_TILE + (long long)threadIdx.x * MT_ITEMS;  // blocked: ranks stay ordered
    Cls c[MT_ITEMS];
    u32 a = 0, b = 0;
#pragma unroll
    for (int i = 0; i < MT_ITEMS; ++i) {
        c[i] = classify(P, first + i);
        a += c[i].isize;
        b += c[i].scope;
    }
    u32 ia = a, ib = b;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u32 ta = __shfl_up_sync(0xffffffffu, ia, off), tb = __shfl_up_sync(0xffffffffu, ib, off);
        if (lane >= off) { ia += ta; ib += tb; }
    }
    if (lane == 31) { s_a[warp] = ia; s_b[warp] = ib; }
    __syncthreads();
    u32 wa = 0, wb = 0;
    for (int w = 0; w < warp; ++w) { wa += s_a[w]; wb += s_b[w]; }
    long long ra = base_isize + pre_isize[blockIdx.x] + wa + ia - a;
    long long rb = base_scope + pre_scope[blockIdx.x] + wb + ib - b;
    u32 mapped = 0;
    int vmax = 0;
#pragma unroll
    for (int i = 0; i < MT_ITEMS; ++i) {
        const int bin = c[i].v < n_bins ? c[i].v : n_bins - 1;
        if ((c[i].isize && ra < MT_CAP) || (c[i].scope && rb < MT_CAP && c[i].cont)) vmax = c[i].v > vmax ? c[i].v : vmax;
        if (c[i].isize) {
            if (ra < MT_CAP) {
                atomicAdd(&hist_isize[bin], 1u);
                if (ra == MT_CAP - 1) out[1] = (u64)(first + i + 1);
            }
            ++ra;
        }
        if (c[i].scope) {
            if (rb < MT_CAP) {
                mapped += c[i].mapped;
                if (c[i].cont) atomicAdd(&hist_cont[bin], 1u);
                if (rb == MT_CAP - 1) out[2] = (u64)(first + i + 1);
            }
            ++rb;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        mapped += __shfl_xor_sync(0xffffffffu, mapped, off);
        const int o = __shfl_xor_sync(0xffffffffu, vmax, off);
        vmax = o > vmax ? o : vmax;
    }
    if (lane == 0 && mapped) atomicAdd(&out[0], (u64)mapped);
    if (lane == 0 && vmax >= n_bins) atomicMax(&out[3], (u64)vmax);
}

// ---- host side: statistics over the histogram ---------------------------------------------

double max_obs_distr(double n, double prob) {  // mathstats normal.MaxObsDistr, restated (A&S 26.2.23)
    const double p = 1 - pow(prob, 1 / n);
    const double q = 1 - p;
    auto rational = [](double t) {
        const double num = (0.010328 * t + 0.802853) * t + 2.515517;
        const double den = ((0.001308 * t + 0.189269) * t + 1.432788) * t + 1.0;
        return t - num / den;
    };
    if (q < 0.5) return -rational(sqrt(-2.0 * log(q)));
    return rational(sqrt(-2.0 * log(1.0 - q)));
}

struct Hist {
    std::vector<u32> h;
    double shift;  // sample value of bin v is v + shift
    double x(size_t v) const { return (double)v + shift; }
    long long n() const {
        long long t = 0;
        for (u32 c : h) t += c;
        return t;
    }
    void mean_sd(double* mean, double* sd) const {
        const double nn = (double)n();
        double s = 0;
        for (size_t v = 0; v < h.size(); ++v)
            if (h[v]) s += (double)h[v] * x(v);
        const double m = s / nn;
        double acc = 0;
        for (size_t v = 0; v < h.size(); ++v)
            if (h[v]) acc += (double)h[v] * ((x(v) * x(v) - 2 * x(v) * m) + m * m);
        *mean = m;
        *sd = sqrt(acc / (nn - 1));
    }
    // AdjustInsertsizeDist (libmetrics.py:22-28): keep mean-k*sd < x < mean+k*sd
    long long trimmed_count(double mean, double sd, double k) const {
        const double lo = mean - k * sd, hi = mean + k * sd;
        long long m = 0;
        for (size_t v = 0; v < h.size(); ++v)
            if (h[v] && x(v) < hi && x(v) > lo) m += h[v];
        return m;
    }
    void trim(double mean, double sd, double k) {
        const double lo = mean - k * sd, hi = mean + k * sd;
        for (size_t v = 0; v < h.size(); ++v)
            if (h[v] && !(x(v) < hi && x(v) > lo)) h[v] = 0;
    }
};

// libmetrics.getdistr (:141-223) over the trimmed histogram
void getdistr_hist(const Hist& H, const int64_t* ref_lengths, int64_t n_refs, besst_libmetrics_out* out,
                   std::vector<double>* adj_out) {
    std::vector<int64_t> all(ref_lengths, ref_lengths + n_refs);
    std::sort(all.begin(), all.end(), [](int64_t a, int64_t b) { return a > b; });
    const size_t nl = (size_t)std::min<int64_t>(n_refs, 1000);
    std::vector<int64_t> largest(all.begin(), all.begin() + nl);
    std::sort(largest.begin(), largest.end());
    size_t vmax = 0;
    for (size_t v = 0; v < H.h.size(); ++v)
        if (H.h[v]) vmax = v;
    const int64_t max_isize = (int64_t)H.x(vmax);
    const int64_t n_bins = max_isize + 1;
    std::vector<double> adj((size_t)n_bins, 0.0);
    int64_t cur_sum = 0;
    for (int64_t l : largest) cur_sum += l;
    int64_t cur_nr = (int64_t)nl;
    const int64_t upper = std::min<int64_t>(n_bins, largest[nl - 1]);
    std::vector<int64_t> tab_nr, tab_sum;
    tab_nr.reserve((size_t)upper + 2);
    tab_sum.reserve((size_t)upper + 2);
    tab_nr.push_back(cur_nr);
    tab_sum.push_back(cur_sum);
    int64_t cur_smallest = largest[0];
    size_t cur_idx = 0;
    for (int64_t isize = 0; isize < upper; ++isize) {
        if (isize > cur_smallest) {
            while (isize > largest[cur_idx]) { ++cur_idx; --cur_nr; cur_sum -= cur_smallest; }  // stale subtraction, as the reference
            tab_nr.push_back(cur_nr);
            tab_sum.push_back(cur_sum);
            cur_smallest = largest[cur_idx];
        } else {
            tab_nr.push_back(cur_nr);
            tab_sum.push_back(cur_sum);
        }
    }
    for (size_t v = 0; v < H.h.size(); ++v) {
        if (!H.h[v]) continue;
        const int64_t obs = (int64_t)H.x(v);
        if (obs > upper) continue;
        const int64_t w0 = tab_sum[(size_t)obs] - (obs - 1) * tab_nr[(size_t)obs];
        const double w = (double)std::max<int64_t>(w0, 10000);
        adj[(size_t)obs] += (double)H.h[v] * (1 / w);
    }
    double tot = 0;
    for (double a : adj) tot += a;
    double cum = 0;
    int64_t cur = 0;
    const double med = tot / 2.0;
    while (cum <= med && cur < n_bins) { cum += adj[(size_t)cur]; ++cur; }
    out->median_adj = cur;
    int64_t modes[21];
    int nm = 0;
    for (int64_t chunk = 1; chunk < 102; chunk += 5) {
        double best = -1;
        int64_t best_i = 0, ci = 0;
        for (int64_t i = 0; i < n_bins; i += chunk, ++ci) {
            double s = 0;
            for (int64_t j = i; j < i + chunk && j < n_bins; ++j) s += adj[(size_t)j];
            if (ci == 0 || s > best) { best = s; best_i = ci; }
        }
        modes[nm++] = (int64_t)(((double)best_i + 0.5) * (double)chunk);
    }
    std::sort(modes, modes + 21);
    out->mode_adj = modes[10];
    double s1 = 0;
    for (int64_t i = 0; i < n_bins; ++i) s1 += (double)i * adj[(size_t)i];
    const double mu = s1 / tot;
    double s2 = 0, s3 = 0;
    for (int64_t i = 0; i < n_bins; ++i) {
        const double d = (double)i - mu;
        s2 += d * d * adj[(size_t)i];
        s3 += d * d * d * adj[(size_t)i];
    }
    const double sigma = sqrt(s2 / tot);
    out->mu_adj = mu;
    out->sigma_adj = sigma;
    out->skew_adj = (s3 / tot) / (sigma * sigma * sigma);
    out->n_bins = n_bins;
    adj_out->swap(adj);
}

}  // namespace

int besst_launch_libmetrics(besst_ctx* ctx, const besst_lib_params& p, const DeviceRecords& rec, const int64_t* ref_lengths,
                            int64_t n_refs, int32_t want_isize, besst_libmetrics_out* out, double* adjusted_distribution,
                            int64_t cap) {
    memset(out, 0, sizeof(*out));
    if (ctx->n_contigs <= 0) { ctx->err = "libmetrics: besst_set_contigs first"; return BESST_E_STATE; }
    int64_t max_len = 0;
    for (int64_t i = 0; i < n_refs; ++i) max_len = std::max(max_len, ref_lengths[i]);
    // |tlen| of a same-contig pair is bounded by the contig length up to aligner slack (soft clips,
    // alignments hanging over the contig end): start with some head-room and re-run with the exact
    // maximum if a sample still falls outside (never clamp: the statistics need exact values)
    int n_bins = (int)std::min<int64_t>(max_len + 1024, (1ll << 28));
    const long long chunk_tiles = (MT_CHUNK + MT_TILE - 1) / MT_TILE;
    std::vector<u32> h_isize, h_cont;
    u64 h_out[4] = {0, 0, 0, 0};
    long long base_isize = 0, base_scope = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        // misc: out[8] hist_isize[n_bins] hist_cont[n_bins] cnt_a[tiles+1] cnt_b[tiles+1]
        const size_t bytes = 4 * (size_t)n_bins * 2 + 4 * (size_t)(chunk_tiles + 1) * 2 + 64;
        BESST_CUDA_TRY(ctx, ctx->misc.ensure(bytes));
        unsigned char* base = ctx->misc.as<unsigned char>();
        u64* d_out = reinterpret_cast<u64*>(base);
        u32* hist_isize = reinterpret_cast<u32*>(base + 64);
        u32* hist_cont = hist_isize + n_bins;
        u32* cnt_a = hist_cont + n_bins;
        u32* cnt_b = cnt_a + chunk_tiles + 1;
        BESST_CUDA_TRY(ctx, cudaMemsetAsync(base, 0, 64 + 4 * (size_t)n_bins * 2, ctx->stream));

        MParams P;
        P.rec = rec;
        P.rows = ctx->rows.as<int4>();
        P.n_contigs = (int)ctx->n_contigs;
        P.orientation = p.orientation; P.min_mapq = p.min_mapq; P.want_isize = want_isize; P.read_len = p.read_len;
        base_isize = 0; base_scope = 0;
        for (long long lo = 0; lo < rec.n; lo += MT_CHUNK) {
            P.lo = lo;
            P.hi = std::min<long long>(rec.n, lo + MT_CHUNK);
            const int n_tiles = (int)((P.hi - P.lo + MT_TILE - 1) / MT_TILE);
            { KTimer kt(ctx, BESST_K_METRICS); k_metrics_count<<<n_tiles, MT_THREADS, 0, ctx->stream>>>(P, cnt_a, cnt_b); }
            { KTimer kt(ctx, BESST_K_METRICS); k_metrics_scan<<<1, 1024, 0, ctx->stream>>>(cnt_a, cnt_b, n_tiles); }
            { KTimer kt(ctx, BESST_K_METRICS); k_metrics_hist<<<n_tiles, MT_THREADS, 0, ctx->stream>>>(P, cnt_a, cnt_b, base_isize, base_scope, hist_isize, hist_cont, n_bins, d_out); }
            BESST_CUDA_TRY(ctx, cudaGetLastError());
            u32 tot[2];
            BESST_CUDA_TRY(ctx, cudaMemcpyAsync(&tot[0], cnt_a + n_tiles, 4, cudaMemcpyDeviceToHost, ctx->stream));
            BESST_CUDA_TRY(ctx, cudaMemcpyAsync(&tot[1], cnt_b + n_tiles, 4, cudaMemcpyDeviceToHost, ctx->stream));
            BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            base_isize += tot[0];
            base_scope += tot[1];
            if ((!want_isize || base_isize >= MT_CAP) && base_scope >= MT_CAP) break;
        }
        h_isize.assign((size_t)n_bins, 0);
        h_cont.assign((size_t)n_bins, 0);
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(h_isize.data(), hist_isize, 4 * (size_t)n_bins, cudaMemcpyDeviceToHost, ctx->stream));
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(h_cont.data(), hist_cont, 4 * (size_t)n_bins, cudaMemcpyDeviceToHost, ctx->stream));
        BESST_CUDA_TRY(ctx, cudaMemcpyAsync(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost, ctx->stream));
        BESST_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if ((long long)h_out[3] < n_bins) break;
        if (attempt == 1 || h_out[3] >= (1ull << 31) - 2) { ctx->err = "libmetrics: |tlen| histogram overflow"; return BESST_E_INVALID; }
        n_bins = (int)h_out[3] + 1;
    }
    {   // records visited by the two capped scans
        long long scanned = 0;
        const long long a = want_isize ? (base_isize >= MT_CAP ? (long long)h_out[1] : rec.n) : 0;
        const long long b = base_scope >= MT_CAP ? (long long)h_out[2] : rec.n;
        scanned = std::max(a, b);
        out->records_scanned = scanned;
    }

    int rc = BESST_OK;
    const bool fr = p.orientation == BESST_ORIENT_FR;
    if (want_isize) {
        Hist H;
        H.h.swap(h_isize);
        H.shift = fr ? 0.0 : 2 * p.read_len;
        long long n = H.n();
        out->n_samples = n;
        if (n <= 1000) {
            rc = 1;  // libmetrics.py:311-314: too few observations (the wrapper exits like the reference)
        } else {
            double mean, sd;
            H.mean_sd(&mean, &sd);
            out->mean_before = mean;
            out->sd_before = sd;
            for (;;) {   // :322-328
                const double k = 1.5 * max_obs_distr((double)n, 0.95);
                const long long m = H.trimmed_count(mean, sd, k);
                const bool removed = m < n;
                H.trim(mean, sd, k);
                n = m;
                H.mean_sd(&mean, &sd);
                if (!removed) break;
            }
            out->n_trimmed = n;
            out->mean_converged = mean;
            out->sd_converged = sd;
            double m3 = 0;
            for (size_t v = 0; v < H.h.size(); ++v)
                if (H.h[v]) { const double d = H.x(v) - mean; m3 += (double)H.h[v] * (d * d * d); }
            m3 /= (double)n;
            out->skewness = m3 / (sd * sd * sd);
            std::vector<double> adj;
            getdistr_hist(H, ref_lengths, n_refs, out, &adj);
            if (adjusted_distribution) {
                const size_t m = (size_t)std::min<int64_t>(out->n_bins, cap);
                memcpy(adjusted_distribution, adj.data(), 8 * m);
            }
        }
    }
    {   // contamination (:86-110)
        Hist H;
        H.h.swap(h_cont);
        H.shift = fr ? 2 * p.read_len : 0.0;
        long long n = H.n();
        double mean = 0, sd = 0;
        out->cont_n_before = n; out->cont_mean_before = 0; out->cont_sd_before = 0;
        if (n > 2) {
            H.mean_sd(&mean, &sd);
            out->cont_mean_before = mean; out->cont_sd_before = sd;
            for (;;) {
                const double k = 1.5 * max_obs_distr((double)n, 0.95);
                const long long m = H.trimmed_count(mean, sd, k);
                const bool removed = m < n;
                if (!(m > 2)) { n = m; break; }
                H.trim(mean, sd, k);
                n = m;
                H.mean_sd(&mean, &sd);
                if (!removed) break;
            }
        }
        out->cont_mapped = (int64_t)h_out[0];
        out->cont_n = n;
        out->cont_mean = mean;
        out->cont_sd = sd;
    }
    return rc;
}
