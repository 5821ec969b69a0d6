/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Plain-C, sequential restatement of the
 * reference's hot path, used as the checker in tests/, in
 * __graft_entry__.smoke() and as bench.py's cpu_baseline.  It is never linked,
 * imported or called by the product path (besst_b200/), which fails loudly if
 * its CUDA library is missing.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * the reference tree).  Integer results are pinned against the reference's own
 * bytecode (oracle/ref_harness.py -> tests/golden/).  Arithmetic that lives in
 * the un-vendored `mathstats==0.2.6.5` (GapEstimator, tr_sk_std_dev,
 * MaxObsDistr, erf) is restated from the published formulas:
 * PARITY UNPINNED at that boundary.
 *
 * Python-3 semantics throughout: true division, round-half-even, insertion-
 * ordered dicts, `**` on floats = libm pow().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/besst_b200.h"

/* ------------------------------------------------------------------------- */
/* mathstats.normaldist.normal (restated; see oracle/mathstats_restated)      */

static double as_erf(double x) {
    const double a1 = 0.254829592, a2 = -0.284496736, a3 = 1.421413741;
    const double a4 = -1.453152027, a5 = 1.061405429, p = 0.3275911;
    double sign = 1.0;
    if (x < 0) sign = -1.0;
    x = fabs(x);
    double t = 1.0 / (1.0 + p * x);
    double y = 1.0 - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * exp(-x * x);
    return sign * y;
}

static double erf_variant(double x, int variant) { return variant == BESST_ERF_LIBM ? erf(x) : as_erf(x); }

static double rational_approximation(double t) {
    const double c0 = 2.515517, c1 = 0.802853, c2 = 0.010328;
    const double d0 = 1.432788, d1 = 0.189269, d2 = 0.001308;
    double numerator = (c2 * t + c1) * t + c0;
    double denominator = ((d2 * t + d1) * t + d0) * t + 1.0;
    return t - numerator / denominator;
}

static double normal_cdf_inverse(double p) {
    if (p < 0.5) return -rational_approximation(sqrt(-2.0 * log(p)));
    return rational_approximation(sqrt(-2.0 * log(1.0 - p)));
}

/* normal.MaxObsDistr (libmetrics.py:23; CreateGraph.py:952,966) */
double besst_oracle_max_obs_distr(double nr_of_obs, double prob) {
    double p = 1 - pow(prob, 1 / nr_of_obs);
    return normal_cdf_inverse(1 - p);
}

/* ------------------------------------------------------------------------- */
/* mathstats...truncatedskewed.param_est (restated)                            */

typedef struct { double g, gp, gb; } gterms;

static gterms g_terms(double d, double mean, double sd, double c_min, double c_max, double r, int variant) {
    double s2 = pow(2.0, 0.5) * sd;
    double A = d + 2 * r - 1 - mean;
    double B = c_min + d + r - mean;
    double C = c_max + d + r - mean;
    double D = c_min + c_max + d + 1 - mean;
    double eA = erf_variant(A / s2, variant), eB = erf_variant(B / s2, variant);
    double eC = erf_variant(C / s2, variant), eD = erf_variant(D / s2, variant);
    double v2 = 2 * pow(sd, 2.0);
    double xA = exp(-pow(A, 2.0) / v2), xB = exp(-pow(B, 2.0) / v2);
    double xC = exp(-pow(C, 2.0) / v2), xD = exp(-pow(D, 2.0) / v2);
    double term1 = (c_min - r + 1) / 2.0 * (eC - eB);
    double term2 = (c_min + c_max + d - mean + 1) / 2.0 * (eD - eC);
    double term3 = (d + 2 * r - mean - 1) / 2.0 * (eA - eB);
    double k = sd / pow(2 * M_PI, 0.5);
    double term4 = k * (xD + xA);
    double term5 = -k * (xC + xB);
    gterms t;
    t.g = term1 + term2 + term3 + term4 + term5;
    t.gp = 0.5 * (eA - eB) + 0.5 * (eD - eC);
    t.gb = (xA - xB - xC + xD) / (pow(2 * M_PI, 0.5) * sd);
    return t;
}

static double py_round_half_even(double x) { return nearbyint(x); /* default FE_TONEAREST = half-even, like Python 3 round() */ }

/* param_est.GapEstimator (call sites CreateGraph.py:537; MakeScaffolds.py:449,453; order_contigs.py:300,308) */
int32_t besst_oracle_gap_estimator(double mean, double sd, double r, double mean_obs, double c1, double c2, int variant) {
    double obs = mean - mean_obs; /* naive gap */
    double c_min = c1 < c2 ? c1 : c2, c_max = c1 < c2 ? c2 : c1;
    double d_upper = (double)(int64_t)(mean + 2 * sd - 2 * r); /* int() truncation */
    double d_lower = (double)(int64_t)(-4 * sd);
    while (d_upper - d_lower > 1) {
        double d_ml = (d_upper + d_lower) / 2.0;
        gterms t = g_terms(d_ml, mean, sd, c_min, c_max, r, variant);
        double aofd = t.gp / t.g; /* IEEE inf/nan where Python would raise */
        double func_of_d = d_ml + aofd * pow(sd, 2.0);
        if (func_of_d > obs) d_upper = d_ml; else d_lower = d_ml;
    }
    double d_ml = (d_upper + d_lower) / 2.0;
    return (int32_t)py_round_half_even(d_ml);
}

/* param_est.tr_sk_std_dev (call site CreateGraph.py:555) */
double besst_oracle_tr_sk_std_dev(double mean, double sd, double r, double c1, double c2, double d, int variant) {
    double c_min = c1 < c2 ? c1 : c2, c_max = c1 < c2 ? c2 : c1;
    gterms t = g_terms(d, mean, sd, c_min, c_max, r, variant);
    double r1 = t.gp / t.g, r2 = t.gb / t.g;
    double e_x = mean - pow(sd, 2.0) * r1;
    double e_x_square = pow(sd, 2.0) + pow(mean, 2.0) + pow(sd, 4.0) * r2 - 2 * mean * pow(sd, 2.0) * r1;
    double e_o = e_x - d;
    double e_o_square = e_x_square - 2 * d * e_x + pow(d, 2.0);
    double var = e_o_square - pow(e_o, 2.0);
    if (!(var >= 0)) return 0.0;
    return pow(var, 0.5);
}

int besst_oracle_gapest_batch(const besst_lib_params* p, const double* mean_obs, const double* len1,
                              const double* len2, int64_t n, int32_t* gap_out, double* sd_out) {
    for (int64_t i = 0; i < n; ++i) {
        int32_t g = besst_oracle_gap_estimator(p->mean_ins_size, p->std_dev_ins_size, p->read_len, mean_obs[i],
                                               len1[i], len2[i], p->erf_variant);
        gap_out[i] = g;
        if (sd_out)
            sd_out[i] = besst_oracle_tr_sk_std_dev(p->mean_ins_size, p->std_dev_ins_size, p->read_len, len1[i],
                                                   len2[i], g, p->erf_variant);
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* mathstats.log_normal_param_est.GapEstimator (restated from the model, see
 * oracle/mathstats_restated/mathstats/log_normal_param_est.py; call sites
 * CreateGraph.py:526, MakeScaffolds.py:425-426, order_contigs.py:304-306)     */

static double ln_Phi(double z) { return 0.5 * erfc(-z / sqrt(2.0)); }
static double ln_F0(double x, double mu, double sigma) { return x > 0 ? ln_Phi((log(x) - mu) / sigma) : 0.0; }
static double ln_F1(double x, double mu, double sigma) {
    return x > 0 ? exp(mu + sigma * sigma / 2.0) * ln_Phi((log(x) - mu - sigma * sigma) / sigma) : 0.0;
}

double besst_oracle_lognormal_g(double d, double mu, double sigma, double c_min, double c_max, double r) {
    double A = d + 2 * r - 1, B = d + c_min + r, C = d + c_max + r, D = d + c_min + c_max + 1;
    double f0A = ln_F0(A, mu, sigma), f0B = ln_F0(B, mu, sigma), f0C = ln_F0(C, mu, sigma), f0D = ln_F0(D, mu, sigma);
    double f1A = ln_F1(A, mu, sigma), f1B = ln_F1(B, mu, sigma), f1C = ln_F1(C, mu, sigma), f1D = ln_F1(D, mu, sigma);
    double piece1 = -(d + 2 * r - 1) * (f0B - f0A) + (f1B - f1A);
    double piece2 = (c_min - r + 1) * (f0C - f0B);
    double piece3 = (d + c_min + c_max + 1) * (f0D - f0C) - (f1D - f1C);
    return piece1 + piece2 + piece3;
}

static double ln_loglik(int64_t d, double mu, double sigma, const int32_t* samples, int64_t n, double c_min, double c_max, double r) {
    double g = besst_oracle_lognormal_g((double)d, mu, sigma, c_min, c_max, r);
    if (!(g > 0)) return -INFINITY;
    double partial[32];
    for (int l = 0; l < 32; ++l) partial[l] = 0.0;
    double v2 = 2.0 * sigma * sigma;
    for (int64_t k = 0; k < n; ++k) {   /* 32 strided partial sums, then a butterfly: the order a warp sums in */
        double lx = log((double)((int64_t)samples[k] + d));
        partial[k & 31] += -lx - (lx - mu) * (lx - mu) / v2;
    }
    for (int off = 16; off > 0; off >>= 1) {
        double q[32];
        for (int l = 0; l < 32; ++l) q[l] = partial[l] + partial[l ^ off];
        for (int l = 0; l < 32; ++l) partial[l] = q[l];
    }
    return partial[0] - (double)n * log(g);
}

int32_t besst_oracle_lognormal_gap_estimator(double mu, double sigma, double r, const int32_t* samples, int64_t n, double c1,
                                             double c2) {
    double c_min = c1 < c2 ? c1 : c2, c_max = c1 < c2 ? c2 : c1;
    int32_t o_min = samples[0];
    for (int64_t k = 1; k < n; ++k) if (samples[k] < o_min) o_min = samples[k];
    double mean_x = exp(mu + sigma * sigma / 2.0);
    double sd_x = sqrt((exp(sigma * sigma) - 1.0) * exp(2.0 * mu + sigma * sigma));
    int64_t lo = (int64_t)(-c_min), alt = 1 - (int64_t)o_min;
    if (alt > lo) lo = alt;
    int64_t hi = (int64_t)(mean_x + 4.0 * sd_x);
    if (hi <= lo) return (int32_t)lo;
    while (hi - lo > 2) {
        int64_t third = (hi - lo) / 3, m1 = lo + third, m2 = hi - third;
        if (ln_loglik(m1, mu, sigma, samples, n, c_min, c_max, r) < ln_loglik(m2, mu, sigma, samples, n, c_min, c_max, r)) lo = m1 + 1;
        else hi = m2 - 1;
    }
    int64_t best = lo;
    double best_l = ln_loglik(lo, mu, sigma, samples, n, c_min, c_max, r);
    for (int64_t d = lo + 1; d <= hi; ++d) {
        double v = ln_loglik(d, mu, sigma, samples, n, c_min, c_max, r);
        if (v > best_l) { best = d; best_l = v; }
    }
    return (int32_t)best;
}

int besst_oracle_gapest_lognormal_batch(double mu, double sigma, double r, const int32_t* samples, const int64_t* row_ptr,
                                        const double* len1, const double* len2, int64_t n, int32_t* gap_out) {
    for (int64_t i = 0; i < n; ++i)
        gap_out[i] = besst_oracle_lognormal_gap_estimator(mu, sigma, r, samples + row_ptr[i], row_ptr[i + 1] - row_ptr[i], len1[i], len2[i]);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* scipy.stats.ks_2samp(...).statistic as used at CreateGraph.py:595           */

static int cmp_double(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

static int64_t upper_bound_d(const double* a, int64_t n, double z) { /* searchsorted(side='right') */
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) / 2;
        if (a[mid] <= z) lo = mid + 1; else hi = mid;
    }
    return lo;
}

double besst_oracle_ks_2samp(const double* d1, int64_t n1, const double* d2, int64_t n2) {
    double* a = (double*)malloc(sizeof(double) * (size_t)(n1 > 0 ? n1 : 1));
    double* b = (double*)malloc(sizeof(double) * (size_t)(n2 > 0 ? n2 : 1));
    memcpy(a, d1, sizeof(double) * (size_t)n1);
    memcpy(b, d2, sizeof(double) * (size_t)n2);
    qsort(a, (size_t)n1, sizeof(double), cmp_double);
    qsort(b, (size_t)n2, sizeof(double), cmp_double);
    double dmax = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        const double* z = pass ? b : a;
        int64_t nz = pass ? n2 : n1;
        for (int64_t i = 0; i < nz; ++i) {
            double cdf1 = (double)upper_bound_d(a, n1, z[i]) / (double)n1;
            double cdf2 = (double)upper_bound_d(b, n2, z[i]) / (double)n2;
            double diff = fabs(cdf1 - cdf2);
            if (diff > dmax) dmax = diff;
        }
    }
    free(a);
    free(b);
    return dmax;
}

/* ------------------------------------------------------------------------- */
/* open-addressing map u64 -> int64 index                                      */

typedef struct { uint64_t* keys; int64_t* vals; int64_t cap, n; } u64map;

static void map_init(u64map* m, int64_t cap) {
    int64_t c = 64;
    while (c < cap * 2) c <<= 1;
    m->cap = c; m->n = 0;
    m->keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)c);
    m->vals = (int64_t*)malloc(sizeof(int64_t) * (size_t)c);
    for (int64_t i = 0; i < c; ++i) m->vals[i] = -1;
}
static void map_free(u64map* m) { free(m->keys); free(m->vals); }
static uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
static int64_t* map_slot(u64map* m, uint64_t key, int* found);
static void map_grow(u64map* m) {
    u64map n;
    map_init(&n, m->cap);
    for (int64_t i = 0; i < m->cap; ++i)
        if (m->vals[i] >= 0) { int f; int64_t* s = map_slot(&n, m->keys[i], &f); *s = m->vals[i]; n.n++; }
    map_free(m);
    *m = n;
}
static int64_t* map_slot(u64map* m, uint64_t key, int* found) {
    uint64_t h = mix64(key) & (uint64_t)(m->cap - 1);
    for (;;) {
        if (m->vals[h] < 0) { m->keys[h] = key; *found = 0; return &m->vals[h]; }
        if (m->keys[h] == key) { *found = 1; return &m->vals[h]; }
        h = (h + 1) & (uint64_t)(m->cap - 1);
    }
}
static int64_t map_get_or_add(u64map* m, uint64_t key, int64_t next_val, int* added) {
    if ((m->n + 1) * 2 > m->cap) map_grow(m);
    int found;
    int64_t* s = map_slot(m, key, &found);
    if (!found) { *s = next_val; m->n++; *added = 1; return next_val; }
    *added = 0;
    return *s;
}
static int64_t map_get(u64map* m, uint64_t key) {
    uint64_t h = mix64(key) & (uint64_t)(m->cap - 1);
    for (;;) {
        if (m->vals[h] < 0) return -1;
        if (m->keys[h] == key) return m->vals[h];
        h = (h + 1) & (uint64_t)(m->cap - 1);
    }
}

/* ------------------------------------------------------------------------- */
/* graph store: one per networkx graph of the reference (G, G_prime)           */

typedef struct {
    u64map map;
    int64_t n_edges, cap_edges;
    uint64_t* key;      /* (u<<32)|v canonical */
    int32_t* nr_links;
    int64_t* obs_sum;
    int64_t* obs_sq;
    int64_t* first_idx;
    /* links in BAM order */
    int64_t n_links, cap_links;
    int64_t* link_edge;
    int32_t* link_ou;
    int32_t* link_ov;
} gstore;

static void gs_init(gstore* g) {
    memset(g, 0, sizeof(*g));
    map_init(&g->map, 1024);
    g->cap_edges = 1024;
    g->key = (uint64_t*)malloc(8 * (size_t)g->cap_edges);
    g->nr_links = (int32_t*)malloc(4 * (size_t)g->cap_edges);
    g->obs_sum = (int64_t*)malloc(8 * (size_t)g->cap_edges);
    g->obs_sq = (int64_t*)malloc(8 * (size_t)g->cap_edges);
    g->first_idx = (int64_t*)malloc(8 * (size_t)g->cap_edges);
    g->cap_links = 4096;
    g->link_edge = (int64_t*)malloc(8 * (size_t)g->cap_links);
    g->link_ou = (int32_t*)malloc(4 * (size_t)g->cap_links);
    g->link_ov = (int32_t*)malloc(4 * (size_t)g->cap_links);
}
static void gs_free(gstore* g) {
    map_free(&g->map);
    free(g->key); free(g->nr_links); free(g->obs_sum); free(g->obs_sq); free(g->first_idx);
    free(g->link_edge); free(g->link_ou); free(g->link_ov);
}

/* the dict-of-dict upsert of CreateEdge, CreateGraph.py:842-862 */
static void gs_add_link(gstore* g, uint32_t n1, uint32_t n2, int32_t obs1, int32_t obs2, int64_t ordinal) {
    uint32_t u = n1, v = n2;
    int32_t ou = obs1, ov = obs2;
    if (u > v) { u = n2; v = n1; ou = obs2; ov = obs1; }
    uint64_t key = ((uint64_t)u << 32) | v;
    int added;
    int64_t e = map_get_or_add(&g->map, key, g->n_edges, &added);
    if (added) {
        if (g->n_edges == g->cap_edges) {
            g->cap_edges *= 2;
            g->key = (uint64_t*)realloc(g->key, 8 * (size_t)g->cap_edges);
            g->nr_links = (int32_t*)realloc(g->nr_links, 4 * (size_t)g->cap_edges);
            g->obs_sum = (int64_t*)realloc(g->obs_sum, 8 * (size_t)g->cap_edges);
            g->obs_sq = (int64_t*)realloc(g->obs_sq, 8 * (size_t)g->cap_edges);
            g->first_idx = (int64_t*)realloc(g->first_idx, 8 * (size_t)g->cap_edges);
        }
        g->key[e] = key; g->nr_links[e] = 0; g->obs_sum[e] = 0; g->obs_sq[e] = 0; g->first_idx[e] = ordinal;
        g->n_edges++;
    }
    int64_t o = (int64_t)obs1 + (int64_t)obs2;
    g->nr_links[e] += 1;
    g->obs_sum[e] += o;
    g->obs_sq[e] += o * o;
    if (g->n_links == g->cap_links) {
        g->cap_links *= 2;
        g->link_edge = (int64_t*)realloc(g->link_edge, 8 * (size_t)g->cap_links);
        g->link_ou = (int32_t*)realloc(g->link_ou, 4 * (size_t)g->cap_links);
        g->link_ov = (int32_t*)realloc(g->link_ov, 4 * (size_t)g->cap_links);
    }
    g->link_edge[g->n_links] = e; g->link_ou[g->n_links] = ou; g->link_ov[g->n_links] = ov;
    g->n_links++;
}

/* ------------------------------------------------------------------------- */
/* PosDirCalculatorPE / PosDirCalculatorMP, CreateGraph.py:1024-1076.          */
/* fwd = read is forward for fr; the strand tests are swapped for rf.          */

static void pos_dir_end(int cont_dir, int read_dir_fwd, int orientation, int64_t cpos, int64_t rpos, int64_t slen,
                        int64_t clen, double read_len, int32_t* obs, int* side_r) {
    int fwd = orientation == BESST_ORIENT_FR ? read_dir_fwd : !read_dir_fwd;
    double o;
    if (cont_dir && fwd) { o = (double)(slen - cpos - rpos); *side_r = 1; }              /* :1025-1027 / :1052-1054 */
    else if (!cont_dir && fwd) { o = (double)(cpos + (clen - rpos)); *side_r = 0; }      /* :1031-1033 / :1058-1060 */
    else if (cont_dir && !fwd) { o = (double)(cpos + rpos) + read_len; *side_r = 0; }    /* :1037-1039 / :1064-1066 */
    else { o = (double)(slen - cpos) - ((double)(clen - rpos) - read_len); *side_r = 1; } /* :1043-1045 / :1070-1072 */
    *obs = (int32_t)o; /* int(): truncation toward zero */
}

/* ------------------------------------------------------------------------- */

typedef struct besst_oracle_result {
    besst_graph_sizes sizes;
    besst_graph_out out; /* arrays owned by this struct */
    besst_link_tuple* tuples; int64_t n_tuples;
    uint64_t* fishy_keys; int32_t* fishy_counts;
    int consistent; /* G_prime's LL edges identical to G's (expected: 1) */
} besst_oracle_result;

typedef struct {
    const besst_lib_params* p;
    int64_t count, non_unique_for_scaf, dups, too_long;
    int32_t prev1, prev2;
    int32_t first1, first2, has_first;   /* the first CreateEdge call of the batch */
} ce_state;

/* CreateEdge, CreateGraph.py:812-871; returns is_dupl */
static int create_edge(ce_state* st, gstore* g, const besst_contig_row* r1, const besst_contig_row* r2, int rev,
                       int mate_rev, int32_t pos, int32_t mpos, int mapq, int64_t* ordinal,
                       besst_link_tuple* tuple_out, int* accepted) {
    const besst_lib_params* p = st->p;
    *accepted = 0;
    if (mapq == 0) st->non_unique_for_scaf += 1;
    int32_t obs1, obs2; int s1, s2;
    pos_dir_end(r1->direction, !rev, p->orientation, r1->position, pos, r1->scaf_length, r1->length, p->read_len, &obs1, &s1);
    pos_dir_end(r2->direction, !mate_rev, p->orientation, r2->position, mpos, r2->scaf_length, r2->length, p->read_len, &obs2, &s2);
    if (!st->has_first) { st->has_first = 1; st->first1 = obs1; st->first2 = obs2; }
    if (obs1 == st->prev1 && obs2 == st->prev2) {
        st->dups += 1;
        if (p->detect_duplicate) return 1;
    }
    if ((double)((int64_t)obs1 + obs2) < p->ins_size_threshold && obs1 > 25 && obs2 > 25) {
        st->count += 1;
        uint32_t n1 = 2u * (uint32_t)r1->scaffold + (uint32_t)s1;
        uint32_t n2 = 2u * (uint32_t)r2->scaffold + (uint32_t)s2;
        gs_add_link(g, n1, n2, obs1, obs2, *ordinal);
        if (tuple_out) {
            if (n1 < n2) { tuple_out->u = n1; tuple_out->v = n2; tuple_out->obs_u = obs1; tuple_out->obs_v = obs2; }
            else { tuple_out->u = n2; tuple_out->v = n1; tuple_out->obs_u = obs2; tuple_out->obs_v = obs1; }
        }
        *accepted = 1;
    } else {
        st->too_long += 1;
    }
    st->prev1 = obs1; st->prev2 = obs2;
    return 0;
}

static const gstore* g_sort_ctx;
static int cmp_edge_by_key(const void* a, const void* b) {
    uint64_t x = g_sort_ctx->key[*(const int64_t*)a], y = g_sort_ctx->key[*(const int64_t*)b];
    return (x > y) - (x < y);
}

/* GiveScoreOnEdges body for one edge, CreateGraph.py:501-614 (normal branch) */
static void score_edge(const besst_lib_params* p, int64_t n, int64_t obs, int64_t obs_sq, double len1, double len2,
                       const int32_t* l1_in, const int32_t* l2_in, int32_t* gap_out, double* score_out,
                       double* ks_out, double* sd_obs_out, double* sd_model_out, uint8_t* flags) {
    double mu = p->mean_ins_size, sigma = p->std_dev_ins_size, r = p->read_len;
    double mean_ = (double)obs / (double)n;                              /* :505 */
    double data_observation = ((double)n * mu - (double)obs) / (double)n; /* :511 */
    int big = (2 * sigma < len1) && (2 * sigma < len2);                   /* :536 */
    double gap;
    if (big) { gap = (double)besst_oracle_gap_estimator(mu, sigma, r, mean_, len1, len2, p->erf_variant); *flags |= BESST_EDGE_BIG; }
    else gap = data_observation;
    *gap_out = (int32_t)gap; /* :541 int() */
    *ks_out = NAN; *sd_obs_out = NAN; *sd_model_out = NAN;
    int neg = (-gap > len1 || -gap > len2);                               /* :542-544: score = 0, the rest is skipped; ks and
                                                                             sd_obs are still reported (diagnostics; the lognormal
                                                                             scoring branch re-derives the verdict from them) */
    double std_dev_d_eq_0 = (big && !neg) ? besst_oracle_tr_sk_std_dev(mu, sigma, r, len1, len2, gap, p->erf_variant)
                                          : 4294967296.0;               /* :548-558 */
    double std_dev;
    if (n - 1 == 0) std_dev = 4294967296.0;                               /* :563-564 */
    else {
        double q = ((double)obs_sq - (double)n * pow(mean_, 2.0)) / (double)(n - 1);
        if (q < 0) { std_dev = NAN; *flags |= BESST_EDGE_CPLX; }
        else std_dev = pow(q, 0.5);                                       /* :561 */
    }
    /* :582-595 */
    double* a = (double*)malloc(sizeof(double) * (size_t)n);
    double* b = (double*)malloc(sizeof(double) * (size_t)n);
    int64_t s1 = 0; int32_t max2 = l2_in[0];
    for (int64_t i = 0; i < n; ++i) { s1 += l1_in[i]; if (l2_in[i] > max2) max2 = l2_in[i]; }
    double l1_mean = (double)s1 / (double)n;
    int64_t s2 = 0;
    for (int64_t i = 0; i < n; ++i) s2 += llabs((int64_t)l2_in[i] - max2);
    double l2_mean = (double)s2 / (double)n;
    for (int64_t i = 0; i < n; ++i) { a[i] = (double)l1_in[i] - l1_mean; b[i] = (double)llabs((int64_t)l2_in[i] - max2) - l2_mean; }
    double ks = besst_oracle_ks_2samp(a, n, b, n);
    free(a); free(b);
    double span_score = n < 5 ? 0.0 : 1 - ks;                             /* :603-606 */
    double std_dev_score;
    if (std_dev_d_eq_0 == 0.0 || std_dev == 0.0 || isnan(std_dev)) std_dev_score = 0.0; /* ZeroDivisionError :610-611 */
    else { double x = std_dev / std_dev_d_eq_0, y = std_dev_d_eq_0 / std_dev; std_dev_score = y < x ? y : x; }
    *ks_out = ks; *sd_obs_out = std_dev;
    if (neg) { *score_out = 0.0; *flags |= BESST_EDGE_NEGGAP; return; }
    *score_out = (std_dev_score > 0.5 && span_score > 0.5) ? std_dev_score + span_score : 0.0; /* :614 */
    *sd_model_out = std_dev_d_eq_0;
}

/* CreateGraph.PE record loop :111-211 + per-edge scoring :498-614 */
besst_oracle_result* besst_oracle_graph_build(const besst_contig_row* rows, int64_t n_contigs, int64_t n_scaffolds,
                                              const besst_lib_params* p, const besst_records* rec) {
    besst_oracle_result* R = (besst_oracle_result*)calloc(1, sizeof(*R));
    gstore G, GP;
    gs_init(&G); gs_init(&GP);
    u64map fishy; map_init(&fishy, 1024);
    int64_t fishy_n = 0, fishy_cap = 1024;
    uint64_t* fishy_key = (uint64_t*)malloc(8 * (size_t)fishy_cap);
    int32_t* fishy_cnt = (int32_t*)malloc(4 * (size_t)fishy_cap);
    int64_t* aligned = (int64_t*)calloc((size_t)(n_contigs > 0 ? n_contigs : 1), 8);
    int64_t tup_cap = 4096; R->tuples = (besst_link_tuple*)malloc(sizeof(besst_link_tuple) * (size_t)tup_cap);
    ce_state st; memset(&st, 0, sizeof(st));
    st.p = p; st.prev1 = p->halo_prev_obs1; st.prev2 = p->halo_prev_obs2; /* counters(0,0,0,0,-1,-1,0) :98 */
    int64_t ctr = 0, non_unique = 0, calls = 0, valid = 0;
    int64_t ord_primary = 0; /* ordinal among accepted primary links */
    int64_t ord_g = 0, ord_gp = 0;
    int scoring = !p->no_score;

    for (int64_t i = 0; i < rec->n; ++i) {
        int32_t tid = rec->tid[i], mtid = rec->mtid[i];
        if (tid < 0 || mtid < 0 || tid >= n_contigs || mtid >= n_contigs) continue;    /* :118-124 */
        const besst_contig_row* r1 = &rows[tid];
        const besst_contig_row* r2 = &rows[mtid];
        if (r1->state == BESST_CTG_ABSENT || r2->state == BESST_CTG_ABSENT) continue;  /* :127-130 */
        valid++;
        int mapq = rec->mapq[i];
        unsigned flag = rec->flag[i];
        int unmapped = (flag & 0x4) != 0, rev = (flag & 0x10) != 0, mrev = (flag & 0x20) != 0;
        int read1 = (flag & 0x40) != 0, read2 = (flag & 0x80) != 0;
        if (mapq >= p->min_mapq || mapq == 0) aligned[tid] += rec->qlen[i];           /* :138-139 */
        if (unmapped && read1 && r1->scaffold != r2->scaffold) {                        /* :141-163 */
            int32_t o1, o2; int s1, s2;                                                 /* CheckDir :678-688 */
            pos_dir_end(r1->direction, !rev, p->orientation, 0, 0, 0, 0, 0.0, &o1, &s1);
            pos_dir_end(r2->direction, !mrev, p->orientation, 0, 0, 0, 0, 0.0, &o2, &s2);
            uint32_t n1 = 2u * (uint32_t)r1->scaffold + (uint32_t)s1, n2 = 2u * (uint32_t)r2->scaffold + (uint32_t)s2;
            uint64_t key = n1 < n2 ? (((uint64_t)n1 << 32) | n2) : (((uint64_t)n2 << 32) | n1);
            int added; int64_t f = map_get_or_add(&fishy, key, fishy_n, &added);
            if (added) {
                if (fishy_n == fishy_cap) { fishy_cap *= 2; fishy_key = (uint64_t*)realloc(fishy_key, 8 * (size_t)fishy_cap); fishy_cnt = (int32_t*)realloc(fishy_cnt, 4 * (size_t)fishy_cap); }
                fishy_key[f] = key; fishy_cnt[f] = 0; fishy_n++;
            }
            fishy_cnt[f] += 1; ctr++;
        }
        if (tid != mtid && mapq == 0) non_unique++;                                     /* :166-167 */
        if (tid != mtid && read2 && !unmapped && mapq >= p->min_mapq) {                 /* :169 */
            int l1 = r1->state == BESST_CTG_LARGE, l2 = r2->state == BESST_CTG_LARGE;
            besst_link_tuple tup; int acc = 0, acc2 = 0;
            if (l1 && l2 && r1->scaffold != r2->scaffold) {                             /* :170-183 */
                calls++;
                int is_dupl;
                if (scoring) is_dupl = create_edge(&st, &G, r1, r2, rev, mrev, rec->pos[i], rec->mpos[i], mapq, &ord_g, &tup, &acc);
                else is_dupl = create_edge(&st, &GP, r1, r2, rev, mrev, rec->pos[i], rec->mpos[i], mapq, &ord_gp, &tup, &acc);
                if (acc) { if (scoring) ord_g++; else ord_gp++; }
                if (p->extend_paths && !is_dupl && scoring) {
                    st.prev1 = -1; st.prev2 = -1;
                    create_edge(&st, &GP, r1, r2, rev, mrev, rec->pos[i], rec->mpos[i], mapq, &ord_gp, NULL, &acc2);
                    if (acc2) ord_gp++;
                }
            } else if (p->extend_paths) {                                               /* :184-206 */
                int v1 = r1->state == BESST_CTG_SMALL, v2 = r2->state == BESST_CTG_SMALL;
                if ((v1 && v2 && r1->scaffold != r2->scaffold) || (v1 && !v2) || (!v1 && v2)) {
                    calls++;
                    create_edge(&st, &GP, r1, r2, rev, mrev, rec->pos[i], rec->mpos[i], mapq, &ord_gp, &tup, &acc);
                    if (acc) ord_gp++;
                }
            }
            if (acc) {
                if (R->n_tuples == tup_cap) { tup_cap *= 2; R->tuples = (besst_link_tuple*)realloc(R->tuples, sizeof(besst_link_tuple) * (size_t)tup_cap); }
                R->tuples[R->n_tuples++] = tup; ord_primary++;
            }
        }
    }

    /* unify the two graphs into the ABI's single edge list:
       scoring+extend: G_prime holds every edge, G = its LL subset;
       scoring only:   G;   no_score: G_prime. */
    gstore* U = (scoring && !p->extend_paths) ? &G : &GP;
    R->consistent = 1;
    int64_t E = U->n_edges;
    int64_t* order = (int64_t*)malloc(8 * (size_t)(E > 0 ? E : 1));
    for (int64_t e = 0; e < E; ++e) order[e] = e;
    g_sort_ctx = U;
    qsort(order, (size_t)E, 8, cmp_edge_by_key);
    int64_t* rank = (int64_t*)malloc(8 * (size_t)(E > 0 ? E : 1));
    for (int64_t k = 0; k < E; ++k) rank[order[k]] = k;

    besst_graph_out* o = &R->out;
    size_t Ez = (size_t)(E > 0 ? E : 1), Lz = (size_t)(U->n_links > 0 ? U->n_links : 1);
    o->edge_u = (uint32_t*)calloc(Ez, 4); o->edge_v = (uint32_t*)calloc(Ez, 4);
    o->nr_links = (int32_t*)calloc(Ez, 4); o->obs_sum = (int64_t*)calloc(Ez, 8); o->obs_sq = (int64_t*)calloc(Ez, 8);
    o->first_idx = (int64_t*)calloc(Ez, 8); o->row_ptr = (int64_t*)calloc(Ez + 1, 8);
    o->gap = (int32_t*)calloc(Ez, 4); o->score = (double*)calloc(Ez, 8); o->ks = (double*)calloc(Ez, 8);
    o->sd_obs = (double*)calloc(Ez, 8); o->sd_model = (double*)calloc(Ez, 8);
    o->fishy = (int32_t*)calloc(Ez, 4); o->flags = (uint8_t*)calloc(Ez, 1);
    o->obs_u = (int32_t*)calloc(Lz, 4); o->obs_v = (int32_t*)calloc(Lz, 4);
    o->aligned_len = aligned;
    for (int64_t k = 0; k < E; ++k) {
        int64_t e = order[k];
        o->edge_u[k] = (uint32_t)(U->key[e] >> 32); o->edge_v[k] = (uint32_t)(U->key[e] & 0xffffffffu);
        o->nr_links[k] = U->nr_links[e]; o->obs_sum[k] = U->obs_sum[e]; o->obs_sq[k] = U->obs_sq[e];
        o->first_idx[k] = U->first_idx[e];
        o->row_ptr[k + 1] = o->row_ptr[k] + U->nr_links[e];
        int64_t f = map_get(&fishy, U->key[e]);
        o->fishy[k] = f >= 0 ? fishy_cnt[f] : 0;
        o->score[k] = NAN; o->ks[k] = NAN; o->sd_obs[k] = NAN; o->sd_model[k] = NAN;
    }
    int64_t* cursor = (int64_t*)malloc(8 * Ez);
    for (int64_t k = 0; k < E; ++k) cursor[k] = o->row_ptr[k];
    for (int64_t l = 0; l < U->n_links; ++l) {
        int64_t k = rank[U->link_edge[l]];
        o->obs_u[cursor[k]] = U->link_ou[l]; o->obs_v[cursor[k]] = U->link_ov[l]; cursor[k]++;
    }
    /* LL flag + cross-check of G against G_prime's LL subset + scoring */
    int64_t n_large2 = 0; /* node ids below 2*n_large are large scaffolds: derive from the rows */
    {
        int32_t max_large = -1;
        for (int64_t c = 0; c < n_contigs; ++c) if (rows[c].state == BESST_CTG_LARGE && rows[c].scaffold > max_large) max_large = rows[c].scaffold;
        n_large2 = 2 * ((int64_t)max_large + 1);
    }
    /* scaffold length by dense index */
    int32_t* slen = (int32_t*)calloc((size_t)(n_scaffolds > 0 ? n_scaffolds : 1), 4);
    for (int64_t c = 0; c < n_contigs; ++c) if (rows[c].state != BESST_CTG_ABSENT) slen[rows[c].scaffold] = rows[c].scaf_length;
    for (int64_t k = 0; k < E; ++k) {
        int ll = (int64_t)o->edge_u[k] < n_large2 && (int64_t)o->edge_v[k] < n_large2;
        if (ll) o->flags[k] |= BESST_EDGE_LL;
        if (ll && scoring) {
            if (U != &G) { /* G must hold the identical edge */
                int64_t ge = map_get(&G.map, ((uint64_t)o->edge_u[k] << 32) | o->edge_v[k]);
                if (ge < 0 || G.nr_links[ge] != o->nr_links[k] || G.obs_sum[ge] != o->obs_sum[k] || G.obs_sq[ge] != o->obs_sq[k]) R->consistent = 0;
            }
            o->flags[k] |= BESST_EDGE_SCORED;
            int64_t b = o->row_ptr[k], n = o->nr_links[k];
            score_edge(p, n, o->obs_sum[k], o->obs_sq[k], slen[o->edge_u[k] >> 1], slen[o->edge_v[k] >> 1],
                       o->obs_u + b, o->obs_v + b, &o->gap[k], &o->score[k], &o->ks[k], &o->sd_obs[k], &o->sd_model[k], &o->flags[k]);
        }
    }
    if (scoring && p->extend_paths) {
        int64_t n_ll = 0;
        for (int64_t k = 0; k < E; ++k) if (o->flags[k] & BESST_EDGE_LL) n_ll++;
        if (n_ll != G.n_edges) R->consistent = 0;
    }
    memset(o->counters, 0, sizeof(o->counters));
    o->counters[BESST_CNT_COUNT] = st.count; o->counters[BESST_CNT_NON_UNIQUE] = non_unique;
    o->counters[BESST_CNT_NON_UNIQUE_SCAF] = st.non_unique_for_scaf; o->counters[BESST_CNT_DUPLICATES] = st.dups;
    o->counters[BESST_CNT_TOO_LONG] = st.too_long; o->counters[BESST_CNT_FISHY] = ctr;
    o->counters[BESST_CNT_CALLS] = calls; o->counters[BESST_CNT_VALID] = valid;
    o->counters[BESST_CNT_LAST_OBS1] = st.prev1; o->counters[BESST_CNT_LAST_OBS2] = st.prev2;
    o->counters[BESST_CNT_FIRST_OBS1] = st.has_first ? st.first1 : 0; o->counters[BESST_CNT_FIRST_OBS2] = st.has_first ? st.first2 : 0;
    R->sizes.n_edges = E; R->sizes.n_links = U->n_links; R->sizes.n_contigs = n_contigs; R->sizes.n_fishy = fishy_n;
    { int64_t nll = 0; for (int64_t k = 0; k < E; ++k) if (o->flags[k] & BESST_EDGE_LL) nll += o->nr_links[k]; R->sizes.n_ll_links = nll; }
    R->fishy_keys = fishy_key; R->fishy_counts = fishy_cnt;
    free(order); free(rank); free(cursor); free(slen);
    map_free(&fishy); gs_free(&G); gs_free(&GP);
    (void)ord_primary;
    return R;
}

void besst_oracle_graph_sizes(const besst_oracle_result* R, besst_graph_sizes* s, int64_t* n_tuples, int* consistent) {
    *s = R->sizes; *n_tuples = R->n_tuples; *consistent = R->consistent;
}

void besst_oracle_graph_fetch(const besst_oracle_result* R, besst_graph_out* out, besst_link_tuple* tuples,
                              uint64_t* fishy_keys, int32_t* fishy_counts) {
    const besst_graph_out* o = &R->out;
    size_t E = (size_t)R->sizes.n_edges, L = (size_t)R->sizes.n_links, C = (size_t)R->sizes.n_contigs;
#define CP(f, n, sz) if (out->f) memcpy(out->f, o->f, (n) * (sz))
    CP(edge_u, E, 4); CP(edge_v, E, 4); CP(nr_links, E, 4); CP(obs_sum, E, 8); CP(obs_sq, E, 8); CP(first_idx, E, 8);
    CP(row_ptr, E + 1, 8); CP(gap, E, 4); CP(score, E, 8); CP(ks, E, 8); CP(sd_obs, E, 8); CP(sd_model, E, 8);
    CP(fishy, E, 4); CP(flags, E, 1); CP(obs_u, L, 4); CP(obs_v, L, 4); CP(aligned_len, C, 8);
#undef CP
    memcpy(out->counters, o->counters, sizeof(o->counters));
    if (tuples) memcpy(tuples, R->tuples, sizeof(besst_link_tuple) * (size_t)R->n_tuples);
    if (fishy_keys) memcpy(fishy_keys, R->fishy_keys, 8 * (size_t)R->sizes.n_fishy);
    if (fishy_counts) memcpy(fishy_counts, R->fishy_counts, 4 * (size_t)R->sizes.n_fishy);
}

void besst_oracle_graph_free(besst_oracle_result* R) {
    besst_graph_out* o = &R->out;
    free(o->edge_u); free(o->edge_v); free(o->nr_links); free(o->obs_sum); free(o->obs_sq); free(o->first_idx);
    free(o->row_ptr); free(o->gap); free(o->score); free(o->ks); free(o->sd_obs); free(o->sd_model); free(o->fishy);
    free(o->flags); free(o->obs_u); free(o->obs_v); free(o->aligned_len);
    free(R->tuples); free(R->fishy_keys); free(R->fishy_counts);
    free(R);
}

/* ------------------------------------------------------------------------- */
/* libmetrics                                                                 */

/* bam_parser.is_proper_aligned_unique_innie / outie, bam_parser.py:22-29 */
static int proper_pair(unsigned flag, int32_t tid, int32_t mtid, int32_t tlen, int mapq, int thr, int innie) {
    int rev = (flag & 0x10) != 0, mrev = (flag & 0x20) != 0, read2 = (flag & 0x80) != 0;
    int mate_unmapped = (flag & 0x8) != 0, secondary = (flag & 0x100) != 0;
    int neg = innie ? tlen < 0 : tlen > 0, posv = innie ? tlen > 0 : tlen < 0;
    int geom = (rev && !mrev && read2 && neg && tid == mtid) || (!rev && mrev && read2 && posv && tid == mtid);
    return geom && !mate_unmapped && (mapq > thr) && !secondary;
}

static void mean_sd(const double* x, int64_t n, double* mean, double* sd) {
    double s = 0;
    for (int64_t i = 0; i < n; ++i) s += x[i];                       /* sum() left to right */
    double m = s / (double)n;
    double acc = 0;                                                  /* libmetrics.py:319 expression */
    for (int64_t i = 0; i < n; ++i) acc += (x[i] * x[i] - 2 * x[i] * m) + pow(m, 2.0);
    *mean = m; *sd = pow(acc / ((double)n - 1), 0.5);
}

/* AdjustInsertsizeDist loop, libmetrics.py:22-28,322-332; min_n guards the
   contamination variant (:100-108: stop when n <= 2) */
static int64_t trim_loop(double* x, int64_t n, double* mean, double* sd, int64_t min_n) {
    for (;;) {
        double k = 1.5 * besst_oracle_max_obs_distr((double)n, 0.95);
        double lo = *mean - k * (*sd), hi = *mean + k * (*sd);
        int64_t m = 0;
        for (int64_t i = 0; i < n; ++i) if (x[i] < hi && x[i] > lo) x[m++] = x[i];
        int removed = m < n;
        if (min_n > 0 && !(m > min_n)) return m; /* :103-108: n_contamine = len(filtered) but mean/sd stay */
        n = m;
        mean_sd(x, n, mean, sd);
        if (!removed) break;
    }
    return n;
}

static int cmp_i64_desc(const void* a, const void* b) { int64_t x = *(const int64_t*)a, y = *(const int64_t*)b; return (y > x) - (y < x); }
static int cmp_i64_asc(const void* a, const void* b) { int64_t x = *(const int64_t*)a, y = *(const int64_t*)b; return (x > y) - (x < y); }

/* libmetrics.getdistr :141-223.  Returns the adjusted distribution (malloc'd). */
static double* getdistr(const double* samples, int64_t n, const int64_t* ref_lengths, int64_t n_refs,
                        besst_libmetrics_out* out) {
    int64_t nl = n_refs < 1000 ? n_refs : 1000;
    int64_t* all = (int64_t*)malloc(8 * (size_t)n_refs);
    memcpy(all, ref_lengths, 8 * (size_t)n_refs);
    qsort(all, (size_t)n_refs, 8, cmp_i64_desc);
    int64_t* largest = (int64_t*)malloc(8 * (size_t)nl);
    memcpy(largest, all, 8 * (size_t)nl);
    qsort(largest, (size_t)nl, 8, cmp_i64_asc);                      /* :142 */
    free(all);
    double mx = samples[0];
    for (int64_t i = 1; i < n; ++i) if (samples[i] > mx) mx = samples[i];
    int64_t max_isize = (int64_t)mx;                                 /* :147 */
    int64_t n_bins = max_isize + 1;
    double* adj = (double*)calloc((size_t)n_bins, 8);
    int64_t cur_sum = 0; for (int64_t i = 0; i < nl; ++i) cur_sum += largest[i];
    int64_t cur_nr = nl;
    int64_t upper_isize = n_bins < largest[nl - 1] ? n_bins : largest[nl - 1]; /* :158 */
    int64_t* tab_nr = (int64_t*)malloc(8 * (size_t)(upper_isize + 2));
    int64_t* tab_sum = (int64_t*)malloc(8 * (size_t)(upper_isize + 2));
    int64_t nt = 0;
    tab_nr[nt] = cur_nr; tab_sum[nt] = cur_sum; nt++;                /* :155 */
    int64_t cur_smallest = largest[0], cur_idx = 0;
    for (int64_t isize = 0; isize < upper_isize; ++isize) {          /* :159-170 */
        if (isize <= cur_smallest) { tab_nr[nt] = cur_nr; tab_sum[nt] = cur_sum; nt++; }
        else {
            while (isize > largest[cur_idx]) { cur_idx++; cur_nr--; cur_sum -= cur_smallest; }
            tab_nr[nt] = cur_nr; tab_sum[nt] = cur_sum; nt++;
            cur_smallest = largest[cur_idx];
        }
    }
    for (int64_t i = 0; i < n; ++i) {                                /* :174-181 */
        int64_t obs = (int64_t)samples[i];
        if (obs > upper_isize) continue;
        int64_t w0 = tab_sum[obs] - (obs - 1) * tab_nr[obs];
        double w = (double)(w0 > 10000 ? w0 : 10000);
        adj[obs] += 1 / w;
    }
    double tot = 0; for (int64_t i = 0; i < n_bins; ++i) tot += adj[i]; /* :189 */
    double cum = 0; int64_t cur = 0; double med = tot / 2.0;
    while (cum <= med) { cum += adj[cur]; cur++; }                   /* :193-197 */
    out->median_adj = cur;
    int64_t modes[21]; int nm = 0;
    for (int64_t chunk = 1; chunk < 102; chunk += 5) {               /* :206-209 */
        double best = -1; int64_t best_i = 0, ci = 0;
        for (int64_t i = 0; i < n_bins; i += chunk, ++ci) {
            double s = 0; for (int64_t j = i; j < i + chunk && j < n_bins; ++j) s += adj[j];
            if (ci == 0 || s > best) { best = s; best_i = ci; }
        }
        modes[nm++] = (int64_t)(((double)best_i + 0.5) * (double)chunk);
    }
    qsort(modes, 21, 8, cmp_i64_asc);
    out->mode_adj = modes[10];                                       /* :211 */
    double s1 = 0; for (int64_t i = 0; i < n_bins; ++i) s1 += (double)i * adj[i];
    double mu = s1 / tot;                                            /* :215 */
    double s2 = 0; for (int64_t i = 0; i < n_bins; ++i) s2 += pow((double)i - mu, 2.0) * adj[i];
    double sigma = sqrt(s2 / tot);                                   /* :216 */
    double s3 = 0; for (int64_t i = 0; i < n_bins; ++i) s3 += pow((double)i - mu, 3.0) * adj[i];
    double m3 = s3 / tot;
    out->mu_adj = mu; out->sigma_adj = sigma; out->skew_adj = m3 / pow(sigma, 3.0); /* :220 */
    out->n_bins = n_bins;
    free(largest); free(tab_nr); free(tab_sum);
    return adj;
}

/* get_metrics sampling :283-356 + get_contamination_metrics :49-131.
   rows[].in_largest marks the 1000 longest references (:231-233). */
int besst_oracle_libmetrics(const besst_contig_row* rows, int64_t n_contigs, const besst_lib_params* p,
                            const besst_records* rec, const int64_t* ref_lengths, int64_t n_refs, int32_t want_isize,
                            besst_libmetrics_out* out, double* adjusted_distribution, int64_t cap) {
    memset(out, 0, sizeof(*out));
    double lib_mean = p->mean_ins_size, lib_sd = p->std_dev_ins_size;
    int64_t scanned = 0;
    if (want_isize) {
        int64_t capn = 1000000, n = 0;
        double* x = (double*)malloc(8 * (size_t)capn);
        int64_t counter = 1;
        for (int64_t i = 0; i < rec->n; ++i) {                        /* :293-303 */
            int32_t tid = rec->tid[i];
            scanned = i + 1;
            int innie = p->orientation == BESST_ORIENT_FR;
            if (proper_pair(rec->flag[i], tid, rec->mtid[i], rec->tlen[i], rec->mapq[i], p->min_mapq, innie)) {
                if (tid >= 0 && tid < n_contigs && rows[tid].in_largest) {
                    double v = (double)llabs((int64_t)rec->tlen[i]);
                    if (!innie) v = v + 2 * p->read_len;
                    x[n++] = v; counter++;
                }
            }
            if (counter > 1000000) break;
        }
        out->n_samples = n;
        if (n <= 1000) { free(x); out->records_scanned = scanned; return 1; } /* :311-314 sys.exit */
        double mean, sd;
        mean_sd(x, n, &mean, &sd);                                   /* :317-319 */
        out->mean_before = mean; out->sd_before = sd;
        n = trim_loop(x, n, &mean, &sd, 0);                          /* :322-332 */
        out->n_trimmed = n; out->mean_converged = mean; out->sd_converged = sd;
        double m3 = 0; for (int64_t i = 0; i < n; ++i) m3 += pow(x[i] - mean, 3.0);
        m3 = m3 / (double)n;
        out->skewness = m3 / pow(sd, 3.0);                           /* :340-341 */
        double* adj = getdistr(x, n, ref_lengths, n_refs, out);      /* :346 */
        if (adjusted_distribution) { int64_t m = out->n_bins < cap ? out->n_bins : cap; memcpy(adjusted_distribution, adj, 8 * (size_t)m); }
        free(adj); free(x);
        lib_mean = out->mu_adj; lib_sd = out->sigma_adj;              /* :355-356 */
    }
    (void)lib_mean; (void)lib_sd;
    /* get_contamination_metrics :49-131 */
    {
        int64_t capn = 1000000, n = 0, sample_counter = 0, counter_total = 0;
        double* x = (double*)malloc(8 * (size_t)capn);
        int64_t i;
        for (i = 0; i < rec->n; ++i) {
            int32_t tid = rec->tid[i];
            if (!(tid >= 0 && tid < n_contigs && rows[tid].in_largest)) continue; /* :65 */
            sample_counter++;
            if (!(rec->flag[i] & 0x4)) counter_total++;              /* :67-68 */
            int want_outie = p->orientation == BESST_ORIENT_FR;
            if (proper_pair(rec->flag[i], tid, rec->mtid[i], rec->tlen[i], rec->mapq[i], p->min_mapq, !want_outie)) {
                double frag = (double)llabs((int64_t)rec->tlen[i]);
                if (want_outie) frag = frag + 2 * p->read_len;       /* :72 / :78 */
                if (p->read_len < frag) x[n++] = frag;
            }
            if (sample_counter >= 1000000) { i++; break; }           /* :83-84 */
        }
        if (i > scanned) scanned = i;
        double mean = 0, sd = 0;
        out->cont_n_before = n; out->cont_mean_before = 0; out->cont_sd_before = 0;
        if (n > 2) {                                                  /* :91-110 */
            mean_sd(x, n, &mean, &sd);
            out->cont_mean_before = mean; out->cont_sd_before = sd;  /* :94-95 */
            n = trim_loop(x, n, &mean, &sd, 2);
        }
        out->cont_mapped = counter_total; out->cont_n = n; out->cont_mean = mean; out->cont_sd = sd;
        free(x);
    }
    out->records_scanned = scanned;
    return 0;
}

int besst_oracle_abi_version(void) { return BESST_ABI_VERSION; }
