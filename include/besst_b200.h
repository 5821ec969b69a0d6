/* besst_b200.h -- C ABI of the B200-native scaffold-graph construction and
 * gap-estimation engine (drop-in for one hot path of ksahlin/BESST).
 *
 * The reference has no FFI on this path; the boundary is defined at its two
 * Python call sites and the mathstats scalar functions, in the
 * style of the reference's one ctypes precedent (BESST/diploid/wrapper_sw.py:12-24
 * <-> BESST/diploid/lib/swmodule.cpp:20-28: ctypes.CDLL, extern "C", caller-
 * allocated out-params, int return code, no exceptions).
 *
 *   entry point              replaces (reference file:line)
 *   -----------------------  -------------------------------------------------
 *   besst_set_contigs        the Contigs/Scaffolds/small_* dict lookups done per
 *                            record in CreateGraph.py:118-130,170-206,819-829
 *   besst_graph_build        CreateGraph.PE record loop :111-211, CreateEdge
 *                            :812-871, PosDirCalculatorPE/MP :1024-1076, CheckDir
 *                            :678-688, fishy counting :141-163, coverage :138-139,
 *                            and the per-edge part of GiveScoreOnEdges :498-614
 *   besst_graph_fetch        (marshalling of the above into caller memory)
 *   besst_libmetrics         libmetrics.get_metrics sampling/trim/getdistr
 *                            :283-356, get_contamination_metrics :49-131,
 *                            bam_parser.is_proper_aligned_unique_innie/outie :22-29
 *   besst_gapest_batch       mathstats param_est.GapEstimator / tr_sk_std_dev at
 *                            CreateGraph.py:537,555; MakeScaffolds.py:449,453;
 *                            order_contigs.py:300,308; pathgaps.py:108,204
 *   besst_libmetrics returns 1 (not an error) when fewer than 1001 insert-size
 *   samples exist (libmetrics.py:311-314).
 *   besst_gapest_lognormal_batch  mathstats log_normal_param_est.GapEstimator (lognormal libraries)
 *   besst_trsk_sd_batch      param_est.tr_sk_std_dev at a caller-given gap
 *                            (CreateGraph.py:555 signature)
 *   besst_graph_view         the same result as views of pinned buffers owned by the ctx
 *   besst_gapest_func_batch  param_est.funcDGeneral, what PreCalcMLvaluesOfdLongContigs
 *                            tabulates (MakeScaffolds.py:68)
 *   besst_links_extract /    the two halves of besst_graph_build on either side of the
 *   besst_links_group /      multi-GPU exchange (no reference counterpart,
 *   besst_runs_route /       the reference is one process): the rank's links are grouped
 *   besst_runs_pack[_peer] / into runs (edge x 2048-link block), whole runs are routed by
 *   besst_runs_to_graph      hash(edge) mod world -- through NCCL or stored straight into
 *                            the destination GPU's peer-mapped buffers -- and the receiver
 *                            merges runs into its share of the CSR
 *   besst_links_partition /  the tuple-level variant of the same exchange (fallback for
 *   besst_links_to_graph     link streams without local order; any caller-made tuple array)
 *   besst_bam_ingest         the pysam.Samfile iteration itself (runBESST:162, libmetrics.py:63,257,293,
 *                            CreateGraph.py:111): BGZF inflate + record decode on the GPU, columns left in HBM
 *
 * Conventions: plain C; all pointers are caller-owned for the duration of the
 * call; nothing is retained after return except inside the ctx; return 0 on
 * success and a negative BESST_E_* code otherwise (message via
 * besst_last_error); never exits or throws.  One ctx per process/GPU; a ctx is
 * not thread-safe.
 */
#ifndef BESST_B200_H
#define BESST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BESST_ABI_VERSION 5

#define BESST_OK 0
#define BESST_E_INVALID -1  /* bad argument */
#define BESST_E_CUDA -2     /* CUDA runtime error (message has the detail) */
#define BESST_E_NOMEM -3
#define BESST_E_STATE -4    /* call order violated (e.g. fetch before build) */
#define BESST_E_NODEVICE -5 /* no usable CUDA device: there is NO CPU fallback */

/* contig states (CreateGraph.py:127: "in Contigs or in small_contigs") */
#define BESST_CTG_ABSENT 0
#define BESST_CTG_LARGE 1 /* in Contigs, scaffold in Scaffolds */
#define BESST_CTG_SMALL 2 /* in small_contigs, scaffold in small_scaffolds */

/* One row per BAM reference id (tid).  32 bytes so a row is two 128-bit loads. */
typedef struct besst_contig_row {
    int32_t state;       /* BESST_CTG_* */
    int32_t scaffold;    /* dense scaffold index: large scaffolds first, in the
                            iteration order of the Scaffolds dict, then small ones */
    int32_t direction;   /* Contig.direction (1 = True) */
    int32_t position;    /* Contig.position */
    int32_t length;      /* Contig.length */
    int32_t scaf_length; /* Scaffold.s_length of its scaffold */
    int32_t in_largest;  /* 1 if tid is among the 1000 longest references
                            (libmetrics.py:231-233) */
    int32_t reserved;
} besst_contig_row;

/* Struct-of-arrays record batch in BAM file order. */
typedef struct besst_records {
    int64_t n;
    const int32_t* tid;   /* pysam rname */
    const int32_t* mtid;  /* mrnm */
    const int32_t* pos;
    const int32_t* mpos;
    const int32_t* tlen;  /* only read by besst_libmetrics */
    const int32_t* qlen;
    const uint16_t* flag;
    const uint8_t* mapq;
    int32_t on_device;    /* 0: host pointers (copied in), 1: device pointers -- the data must be complete with
                             respect to the ctx's stream: synchronise the producing stream first, or make the
                             ctx run on it (besst_set_stream) */
    int32_t reserved;
    const uint32_t* packed; /* optional: flag | mapq << 12 | qlen << 20 in ONE column (BESST_PACK_RECORD).  When given, the
                               graph build reads it INSTEAD of flag / mapq / qlen: 20 instead of 23 bytes per record over
                               PCIe and out of HBM (the ingest library writes it while decoding).  Needs flag < 4096 (all
                               SAM flag bits) and qlen < 4096; besst_libmetrics still reads the three plain columns. */
} besst_records;
#define BESST_PACK_RECORD(flag, mapq, qlen) (((uint32_t)(flag) & 0xfffu) | ((uint32_t)(mapq) << 12) | ((uint32_t)(qlen) << 20))

#define BESST_ORIENT_FR 0
#define BESST_ORIENT_RF 1
#define BESST_ERF_AS7126 0 /* Abramowitz-Stegun 7.1.26, mathstats' erf */
#define BESST_ERF_LIBM 1

typedef struct besst_lib_params {
    int32_t orientation;        /* BESST_ORIENT_* (param.orientation) */
    int32_t min_mapq;           /* param.min_mapq */
    int32_t detect_duplicate;   /* param.detect_duplicate */
    int32_t extend_paths;       /* param.extend_paths */
    int32_t no_score;           /* param.no_score */
    int32_t erf_variant;        /* BESST_ERF_* */
    double read_len;            /* param.read_len (may be non-integral) */
    double mean_ins_size;       /* param.mean_ins_size */
    double std_dev_ins_size;    /* param.std_dev_ins_size */
    double ins_size_threshold;  /* param.ins_size_threshold */
    /* multi-GPU halo: the (obs1,obs2) of the last CreateEdge call
       made by any preceding rank, (-1,-1) for rank 0 / single GPU */
    int32_t halo_prev_obs1;
    int32_t halo_prev_obs2;
} besst_lib_params;

/* counters[] slots (Parameter.py:113-124 `counters`, CreateGraph.py:98-100) */
#define BESST_CNT_COUNT 0           /* counter.count */
#define BESST_CNT_NON_UNIQUE 1      /* counter.non_unique */
#define BESST_CNT_NON_UNIQUE_SCAF 2 /* counter.non_unique_for_scaf */
#define BESST_CNT_DUPLICATES 3      /* counter.nr_of_duplicates */
#define BESST_CNT_TOO_LONG 4        /* counter.reads_with_too_long_insert */
#define BESST_CNT_FISHY 5           /* ctr ("NR OF FISHY READ LINKS") */
#define BESST_CNT_CALLS 6           /* records that reached CreateEdge */
#define BESST_CNT_VALID 7           /* records with both contigs present */
#define BESST_CNT_LAST_OBS1 8       /* (obs1,obs2) of the last CreateEdge call, */
#define BESST_CNT_LAST_OBS2 9       /*   -1,-1 if none: the next rank's halo     */
#define BESST_CNT_FIRST_OBS1 10     /* (obs1,obs2) of the FIRST CreateEdge call (valid if CALLS > 0): lets the  */
#define BESST_CNT_FIRST_OBS2 11     /*   multi-GPU driver check a slice against its halo without a tail pass   */
#define BESST_CNT_POS_TILES 12      /* 128-record tiles whose pos / mpos columns the record kernel fetched (its real input: tiles
                                       without a CreateEdge candidate are classified from tid / mtid / flags alone) */
#define BESST_N_COUNTERS 16

/* edge flags */
#define BESST_EDGE_LL 1      /* both scaffolds large: edge of G (and of G_prime when extend_paths) */
#define BESST_EDGE_SCORED 2  /* gap/score computed (LL edge and not no_score) */
#define BESST_EDGE_NEGGAP 4  /* -gap > len: score = 0 and the per-scaffold lists are kept (CreateGraph.py:542-544) */
#define BESST_EDGE_BIG 8     /* 2*sigma < len1 and 2*sigma < len2: GapEstimator used (:536) */
#define BESST_EDGE_CPLX 16   /* (obs_sq - n*mean^2) < 0: reference would produce a complex sd (:561) */

typedef struct besst_graph_sizes {
    int64_t n_edges;   /* E: distinct link edges */
    int64_t n_links;   /* Lk: accepted links = sum of nr_links */
    int64_t n_contigs; /* C */
    int64_t n_fishy;   /* fishy keys (unmapped-read1 records, :141-163) seen by this build */
    int64_t n_ll_links; /* links on large-large edges (the ones that are scored) */
} besst_graph_sizes;

/* Caller-allocated result arrays (host).  Edges are sorted by (edge_u, edge_v),
 * edge_u < edge_v, node id = 2*scaffold + (side == 'R').  obs_u/obs_v are the
 * per-scaffold observation lists (CreateGraph.py:848-849,853-854) in BAM order
 * within each edge, obs_u on edge_u's scaffold; `observations` = obs_u+obs_v. */
typedef struct besst_graph_out {
    uint32_t* edge_u;    /* [E] */
    uint32_t* edge_v;    /* [E] */
    int32_t* nr_links;   /* [E] */
    int64_t* obs_sum;    /* [E] 'obs' */
    int64_t* obs_sq;     /* [E] 'obs_sq' */
    int64_t* first_idx;  /* [E] ordinal (among accepted links, BAM order) of the
                            edge's first link: edge insertion order */
    int64_t* row_ptr;    /* [E+1] */
    int32_t* gap;        /* [E] int(gap), valid if flags & SCORED */
    double* score;       /* [E] valid if flags & SCORED */
    double* ks;          /* [E] KS statistic (diagnostic), valid if SCORED and not NEGGAP */
    double* sd_obs;      /* [E] sample sd (:561) */
    double* sd_model;    /* [E] tr_sk_std_dev or 2**32 (:548-558) */
    int32_t* fishy;      /* [E] fishy_edges count for this node pair (:161-162) */
    uint8_t* flags;      /* [E] BESST_EDGE_* */
    int32_t* obs_u;      /* [Lk] */
    int32_t* obs_v;      /* [Lk] */
    int64_t* aligned_len; /* [C] cont_aligned_len[contig][0] (:139) */
    int64_t counters[BESST_N_COUNTERS];
} besst_graph_out;

/* Link tuple stream between the two halves (multi-GPU exchange unit, 16 B). */
typedef struct besst_link_tuple {
    uint32_t u;     /* canonical: u < v */
    uint32_t v;
    int32_t obs_u;
    int32_t obs_v;
} besst_link_tuple;

typedef struct besst_libmetrics_out {
    /* insert-size estimate (libmetrics.py:283-356); valid if want_isize */
    int64_t n_samples;        /* len(ins_size_reads) before filtering */
    int64_t n_trimmed;        /* after the AdjustInsertsizeDist loop */
    double mean_before, sd_before;   /* :318-319 */
    double mean_converged, sd_converged; /* :331-332 */
    double skewness;          /* :340-341 */
    double mu_adj, sigma_adj, skew_adj; /* getdistr :215-220 */
    int64_t median_adj, mode_adj;       /* getdistr :189-211 */
    int64_t n_bins;           /* len(adjusted_distribution) = max_isize+1 */
    /* contamination (libmetrics.py:49-131) */
    int64_t cont_mapped;      /* counter_total */
    int64_t cont_n;           /* n_contamine after trim */
    double cont_mean, cont_sd;
    int64_t records_scanned;  /* records visited by the capped scans */
    int64_t cont_n_before;    /* len(contamination_reads) before the trim loop (:89) */
    double cont_mean_before, cont_sd_before;   /* :92-95, printed to Information; 0 unless cont_n_before > 2 */
} besst_libmetrics_out;

typedef struct besst_ctx besst_ctx;

int besst_abi_version(void);
/* device < 0: use the current CUDA device */
besst_ctx* besst_create(int device);
void besst_destroy(besst_ctx* ctx);
const char* besst_last_error(besst_ctx* ctx);

int besst_set_contigs(besst_ctx* ctx, const besst_contig_row* rows, int64_t n_contigs,
                      int64_t n_scaffolds, int64_t n_large_scaffolds);

/* Several contig tables resident in HBM at once: a scaffolding run walks its libraries in sequence
 * (runBESST:143-231) and every library sees a different Contigs/Scaffolds state.  besst_contigs_select
 * makes `slot` (0 .. BESST_MAX_TABLES-1) the table besst_set_contigs writes and the builds read; slot 0
 * is selected at creation.  Switching is O(1): no copy, no synchronisation. */
#define BESST_MAX_TABLES 8
int besst_contigs_select(besst_ctx* ctx, int32_t slot);

/* records -> CSR edges + per-edge statistics + gap/score, resident in HBM. */
int besst_graph_build(besst_ctx* ctx, const besst_lib_params* params,
                      const besst_records* records, besst_graph_sizes* sizes);
/* copy the last build's result into caller-allocated host arrays */
int besst_graph_fetch(besst_ctx* ctx, besst_graph_out* out);

/* the same result as zero-copy views: the arrays of `out` are set to PINNED host buffers owned by
 * the ctx (filled by this call); they stay valid until the next besst_graph_view / besst_destroy on
 * this ctx.  Avoids the page faults and the bounce buffer of a pageable destination (0.6 GB at
 * config 3). */
int besst_graph_view(besst_ctx* ctx, besst_graph_out* out);

/* multi-GPU halves: extract leaves the accepted link tuples (BAM order) in HBM
 * and reports how many; tuples_device returns the device pointer for the NCCL
 * exchange; links_to_graph consumes an (exchanged) device tuple array. */
int besst_links_extract(besst_ctx* ctx, const besst_lib_params* params,
                        const besst_records* records, int64_t* n_tuples);
int besst_links_tuples_device(besst_ctx* ctx, const besst_link_tuple** tuples, int64_t* n_tuples);
int besst_links_fishy_device(besst_ctx* ctx, const uint64_t** keys, int64_t* n_keys);
int besst_links_partials(besst_ctx* ctx, int64_t* aligned_len_host /*[C]*/, int64_t* counters_host /*[16]*/);
/* device views of the same partial sums (aligned_len[C], counters[16], int64) for an in-place
 * NCCL all-reduce; valid until the next extract.  The two live in ONE allocation, counters_device ==
 * aligned_len_device + C + 1: a single all-reduce over C + 1 + 16 words covers both (zero the per-rank
 * slots BESST_CNT_LAST_OBS1 .. BESST_CNT_FIRST_OBS2 first). */
int besst_links_partials_device(besst_ctx* ctx, int64_t** aligned_len_device, int64_t** counters_device);
/* host copies of the extracted tuple stream / fishy keys (tests, debugging); either may be NULL */
int besst_links_fetch(besst_ctx* ctx, besst_link_tuple* tuples_host, uint64_t* fishy_keys_host);
/* stable partition of the extracted tuples and fishy keys into `world` destination buckets
 * (bucket d = hash(u,v) mod world, BAM order kept inside a bucket), written to caller-provided
 * DEVICE buffers of n_tuples / n_fishy_keys elements; out_ordinals_device (optional, n_tuples
 * uint32) receives each bucketed tuple's ordinal in this rank's BAM-ordered stream, from which the
 * receiver rebuilds the global first-appearance order of its edges; *_counts[world] (host) receive
 * the bucket sizes.  world <= 16.  out_tuples_device == NULL: only the fishy keys are partitioned (the
 * links travel as runs, see besst_links_group). */
int besst_links_partition(besst_ctx* ctx, int32_t world, besst_link_tuple* out_tuples_device,
                          uint32_t* out_ordinals_device, uint64_t* out_fishy_device, int64_t* tuple_counts,
                          int64_t* fishy_counts);
int besst_links_to_graph(besst_ctx* ctx, const besst_lib_params* params,
                         const besst_link_tuple* tuples_device, int64_t n_tuples,
                         const uint64_t* fishy_keys_device, int64_t n_fishy_keys,
                         besst_graph_sizes* sizes);

/* ---- run-level multi-GPU exchange (preferred over the tuple-level one) -------------------------------
 * After besst_links_extract, besst_links_group groups the rank's accepted links by edge inside blocks
 * of 2048 consecutive links (BAM order kept) and leaves one RUN per (block, edge).  Whole runs are
 * routed by hash(u,v) mod world: 8 bytes per link (obs_u, obs_v) plus one 24-byte descriptor per run
 * cross NVLink instead of 20 bytes per link, and the receiver starts at the run merge -- it never
 * regroups links.  returns 1 (not an error) when the stream has no local order: use the tuple path. */
typedef struct besst_run_desc {
    uint32_t u, v;    /* edge, u < v */
    uint32_t count;   /* links of the run */
    uint32_t first;   /* ordinal of its first link in the SOURCE rank's accepted-link stream (BAM order) */
    uint32_t offset;  /* its first observation inside the (source -> destination) observation segment */
    uint32_t block;   /* block of the source's stream it comes from: BAM order among the runs of one edge */
} besst_run_desc;
int besst_links_group(besst_ctx* ctx, int64_t* n_runs);
/* besst_links_group + besst_runs_route + the fishy-key half of besst_links_partition queued back to back with ONE host
 * read for all the sizes a rank contributes to the exchange's count matrix (a host round trip costs as much as the small
 * kernels in between).  summary[8] = { runs usable (0: the stream has no local order, take the tuple path), runs,
 * CreateEdge calls, last obs1, last obs2, first obs1, first obs2, accepted links }; *_counts[world] per destination;
 * out_fishy_device (n_fishy keys, may be NULL when there are none) receives the keys bucketed by destination. */
int besst_exchange_prepare(besst_ctx* ctx, int32_t world, uint64_t* out_fishy_device, int64_t* summary,
                           int64_t* link_counts, int64_t* run_counts, int64_t* fishy_counts);
/* bytes per link of the exchanged observations: 4 (obs_u | obs_v << 16) when 0 < ins_size_threshold <= 65535
 * -- every accepted observation is below the threshold, CreateGraph.py:840 -- else 8 (two int32) */
int besst_runs_obs_bytes(const besst_lib_params* params);
/* per destination: links and runs this rank will send (host arrays of `world` entries) */
int besst_runs_route(besst_ctx* ctx, int32_t world, int64_t* link_counts, int64_t* run_counts);
/* fill caller-provided DEVICE buffers, destination-major: out_obs = the observations of sum(link_counts) links
 * (besst_runs_obs_bytes each), out_desc = sum(run_counts) descriptors */
int besst_runs_pack(besst_ctx* ctx, int32_t world, int32_t* out_obs_device, besst_run_desc* out_desc_device);
/* the same, fused with the exchange: obs_ptrs[d] / desc_ptrs[d] (host arrays of `world` DEVICE pointers) address
 * the start of this rank's segment inside destination d's receive buffers -- peer-mapped memory of GPU d
 * (CUDA IPC / symmetric memory over NVLink).  The kernel's stores ARE the all-to-all; the caller brackets
 * it with a cross-GPU barrier on each side. */
int besst_runs_pack_peer(besst_ctx* ctx, int32_t world, int32_t* const* obs_ptrs, besst_run_desc* const* desc_ptrs);
/* build this rank's share of the graph from the runs received from all sources (source-major device
 * buffers).  src_*_counts[world]: what each source sent here; src_first_base[world]: number of accepted
 * links on all ranks before the source (first_idx becomes a GLOBAL ordinal); block_bits: bits of the
 * largest block index on any rank. */
int besst_runs_to_graph(besst_ctx* ctx, const besst_lib_params* params, const int32_t* obs_device, int64_t n_links,
                        const besst_run_desc* desc_device, int64_t n_runs, int32_t world, int32_t block_bits,
                        const int64_t* src_run_counts, const int64_t* src_link_counts, const int64_t* src_first_base,
                        const uint64_t* fishy_keys_device, int64_t n_fishy_keys, besst_graph_sizes* sizes);

/* library metrics: capped BAM-order sampling + histogram on the GPU, O(bins)
 * statistics on the host side of the shim.  lengths = BAM header lengths.
 * adjusted_distribution (optional, may be NULL) receives min(n_bins, cap) bins. */
int besst_libmetrics(besst_ctx* ctx, const besst_lib_params* params, const besst_records* records,
                     const int64_t* ref_lengths, int64_t n_refs, int32_t want_isize,
                     besst_libmetrics_out* out, double* adjusted_distribution, int64_t cap);

/* batched GapEstimator + tr_sk_std_dev (host arrays in, host arrays out).  Contig lengths are fp64
 * like every other argument of the mathstats functions: MakeScaffolds.py:68 passes c1 = c2 =
 * mean + 4*stdDev of the ESTIMATED library parameters, which is never integral. */
int besst_gapest_batch(besst_ctx* ctx, const besst_lib_params* params, const double* mean_obs,
                       const double* len1, const double* len2, int64_t n,
                       int32_t* gap_out, double* sd_out);

/* d[i] + sigma^2 g'(d[i])/g(d[i]) for contig lengths len1[i], len2[i]: the left-hand side of the ML equation
 * (mathstats funcDGeneral), what PreCalcMLvaluesOfdLongContigs tabulates (MakeScaffolds.py:68) */
int besst_gapest_func_batch(besst_ctx* ctx, const besst_lib_params* params, const double* d, const double* len1,
                            const double* len2, int64_t n, double* func_out);

/* lognormal GapEstimator (mathstats.log_normal_param_est.GapEstimator; CreateGraph.py:526, MakeScaffolds.py:425-426,
 * order_contigs.py:304-306): the ML gap of every edge from its RAW observations (the `observations` payload of
 * the CSR: samples[row_ptr[i] .. row_ptr[i+1])), one warp per edge.  mu_ln / sigma_ln = param.lognormal_mean /
 * param.lognormal_sigma (libmetrics.py:385-388).  Host arrays in and out. */
int besst_gapest_lognormal_batch(besst_ctx* ctx, double mu_ln, double sigma_ln, double read_len, const int32_t* samples,
                                 const int64_t* row_ptr, const double* len1, const double* len2, int64_t n, int32_t* gap_out);

/* tr_sk_std_dev(mean, sigma, read_len, len1[i], len2[i], gap[i]) (host arrays in/out) */
int besst_trsk_sd_batch(besst_ctx* ctx, const besst_lib_params* params, const double* gap, const double* len1,
                        const double* len2, int64_t n, double* sd_out);

/* Order-dependent pruning of G_prime in high-density regions (remove_edges_below_threshold, CreateGraph.py:355-374)
 * on the CSR edge list, host side: the weak edges (nr_links < expected_links) arrive in the order networkx's
 * G.edges() would yield them; an edge is dropped iff BOTH endpoints still have more than `min_neighbours`
 * neighbours at that moment, and dropping it lowers both degrees.  degree[] (per node id, in/out) counts the
 * intra-scaffold edge as well.  dropped[i] = 1 for the removed edges.  Returns how many were removed.
 * No ctx, no device: a sequential O(n_weak) loop that does not belong in Python at 1e6 edges. */
int64_t besst_csr_prune_dense(int64_t n_weak, const uint32_t* weak_u, const uint32_t* weak_v, int32_t* degree,
                              int32_t min_neighbours, uint8_t* dropped);

/* ---- BAM ingest on the device ----------------------------------------------------------------
 * Sorted BAM file -> the record columns of besst_records RESIDENT IN HBM.  The file crosses PCIe compressed (windows of
 * BGZF blocks, double-buffered against the kernels); every BGZF block is inflated by one warp, its CRC-32 checked, record
 * boundaries are found per block and verified on the host in O(blocks), the fixed-core fields and the CIGAR-derived
 * lengths (pysam 0.8.4's qlen) are decoded straight into the columns.  `out` receives DEVICE pointers
 * (on_device = 1; packed is NULL when a record does not fit the packed column) owned by the ctx and valid until the next
 * besst_bam_ingest / besst_destroy: hand it to besst_libmetrics / besst_graph_build as is -- no record visits host memory.
 * Handles records that straddle BGZF blocks and windows (any BGZF writer, not only htslib's).  There is no host fallback:
 * without a device the call fails like every other entry point; the host-thread ingest is libbesst_bamio.so. */
#define BESST_BAM_NO_CRC 1       /* flags: skip the CRC-32 check of the inflated blocks */
#define BESST_BAM_BLIND_SEEDS 2  /* flags (testing): seed every block's record hop at its first byte, so that the host
                                    verification has to repair every block that starts inside a record */
typedef struct besst_bam_ingest_stats {
    int64_t compressed_bytes, uncompressed_bytes, blocks, records, windows;
    int64_t rescans;          /* blocks whose seed the verification rejected and re-hopped */
    double seconds_total;     /* host wall clock of the call */
    double seconds_read;      /* of which: file -> pinned staging (host threads) */
    float ms_inflate, ms_scan, ms_decode;   /* device time (CUDA events), summed over the windows */
    int32_t crc_checked;
} besst_bam_ingest_stats;
int besst_bam_ingest(besst_ctx* ctx, const char* path, int64_t head_records, int32_t flags, besst_records* out,
                     besst_bam_ingest_stats* stats /* may be NULL */);
/* One PART of the file, for a multi-GPU ingest (rank `part` of `n_parts`, one ctx / GPU each): with D = the file offset of
 * the BGZF block in which the header ends, the part owns the blocks that start inside
 * [D + (size - D) part / n_parts, D + (size - D) (part + 1) / n_parts) and the records that start in them; its last record
 * may end in the next part's first blocks (up to 1 MB of them are inflated along, BESST_BAM_TAIL).  Every part reads the
 * header itself.  BGZF virtual offsets (block file offset << 16 | offset in the inflated block) tie the parts together:
 * *landing_voffset = the first record behind the part, *first_voffset = the part's first record (-1 / -1: no record
 * starts in the part).  A part behind the header finds its first record by a plausibility test that nothing inside the
 * part can verify: the caller compares first_voffset[r] with landing_voffset[r - 1] (one all_gather) and repeats a
 * mismatching part with start_voffset = landing_voffset[r - 1] (-1: not known).  The concatenation of the parts' columns
 * in part order is the whole file's.  besst_bam_ingest is part 0 of 1. */
int besst_bam_ingest_part(besst_ctx* ctx, const char* path, int64_t head_records, int32_t flags, int32_t part, int32_t n_parts,
                          int64_t start_voffset, besst_records* out, besst_bam_ingest_stats* stats /* may be NULL */,
                          int64_t* first_voffset /* may be NULL */, int64_t* landing_voffset /* may be NULL */);
/* header of the last ingested file, and rlen (l_seq) / alen (reference span) of its first head_records records (what
 * libmetrics.py:246-266 reads); besst_bam_ingest_head returns how many were written (<= cap) */
int64_t besst_bam_ingest_n_refs(besst_ctx* ctx);
const char* besst_bam_ingest_ref_name(besst_ctx* ctx, int64_t i);
int64_t besst_bam_ingest_ref_length(besst_ctx* ctx, int64_t i);
int64_t besst_bam_ingest_head(besst_ctx* ctx, int32_t* rlen, int32_t* alen, int64_t cap);
/* device -> host copy on the ctx's stream (tests and debugging: read back columns the ctx owns) */
int besst_device_read(besst_ctx* ctx, const void* device_ptr, void* host_ptr, int64_t bytes);

/* ---- path search between scaffolds ---------------------------------------------------------------
 * ELS.BetweenScaffolds (ExtendLargeScaffolds.py:665-712) on a CSR rendering of G_prime: for every start node the
 * reference's default traversal (find_all_paths_for_start_node_DFS_dynamic_programming_ish, :526-663: best-first over a heap
 * of (nr_links, node, path), one bad-neighbour set per search, head_dict pruning, the --iter cap) and ScorePaths (:28-133)
 * for the paths it finds.  Node id = 2 * rank(scaffold key) + (side == 'R') with ranks ascending in the key, so that
 * integer order is the order of the reference's (scaffold, side) tuples; the contig edge id <-> id ^ 1 carries
 * adj_links = -1.  `order` = the start nodes in the order the reference would pop them from iter_nodes (the searches are
 * independent given that order: already_visited = the earlier start nodes, end = is_end minus the start nodes up to the
 * current one) -- they run on n_threads host threads (<= 0: all cores).  No ctx, no device: an irregular best-first walk
 * per start node.  The result keeps the paths that pass ScorePaths' filter (score >= score_cutoff, len > 2 unless
 * no_score), in start order then in the order found, with the good / bad link weights of the score (the caller forms
 * good / bad with the reference's arithmetic; the contamination variant halves good first).  NULL on bad arguments. */
typedef struct besst_paths besst_paths;
besst_paths* besst_paths_between(int64_t n_nodes, const int64_t* adj_ptr, const int32_t* adj_node, const int32_t* adj_links,
                                 const uint8_t* is_end, const int32_t* order, int64_t n_order, int64_t path_threshold,
                                 double score_cutoff, int32_t no_score, int32_t contamination, int32_t n_threads);
int64_t besst_paths_count(const besst_paths* p);
int32_t besst_paths_hit_threshold(const besst_paths* p);   /* param.hit_path_threshold (:561) */
int64_t besst_paths_pops(const besst_paths* p);            /* heap pops over all searches (work done) */
int besst_paths_arrays(const besst_paths* p, const int64_t** path_ptr /*[n + 1]*/, const int32_t** nodes, const int64_t** good,
                       const int64_t** bad, const int32_t** start_index);
void besst_paths_free(besst_paths* p);

/* RemoveAmbiguousRegionsUsingScore (MakeScaffolds.py:206-240, per-node rule remove_edges :156-204) on the scored link edges
 * of G, host side: sequential and order dependent like besst_csr_prune_dense.  The edges arrive in the order G.edges()
 * yields them; order[] = their indices sorted by score, descending and stable (the reference's processing order); node ids
 * preserve the order of the (scaffold, side) tuples.  At each endpoint of each edge in turn: the zero-score edges go, of the
 * others all but the best go -- or all of them when the two best are within a factor 0.8 (logged: amb_best / amb_second
 * receive the edge indices of every such event, capacity 2 * n_edges).  removed[e] = 1 for the edges to drop from G (and
 * from G_prime when extend_paths).  Returns the number of ambivalent events, -1 on bad arguments. */
int64_t besst_scaffold_prune_ambiguous(int64_t n_nodes, int64_t n_edges, const int32_t* eu, const int32_t* ev, const double* score,
                                       const int64_t* order, uint8_t* removed, int64_t* amb_best, int64_t* amb_second);

/* run all work of this ctx on a caller-owned CUDA stream (a cudaStream_t passed as void*; NULL
 * restores the ctx's own non-blocking stream; pass cudaStreamLegacy (0x1) for the legacy default stream).  Lets a host framework order the library's kernels with its own
 * work (NCCL collectives, CUDA-event timing) without device-wide synchronisation. */
int besst_set_stream(besst_ctx* ctx, void* cuda_stream);

/* device-side timing of the last besst_graph_build, CUDA events on the
 * library's stream: total and per-stage milliseconds */
#define BESST_N_STAGES 8
int besst_last_timing(besst_ctx* ctx, float* total_ms, float* stage_ms /*[BESST_N_STAGES]*/);
int besst_kernel_launches(besst_ctx* ctx, int64_t* n_launches);

/* per-kernel CUDA-event timing (for bench.py's roofline): when enabled every
 * kernel launch of a build is bracketed by an event pair on the library's
 * stream; besst_kernel_profile returns (kernel id, milliseconds) per launch of
 * the last build, in launch order.  Returns the number of launches written. */
#define BESST_K_EXTRACT 0      /* k_extract_links          (K1) */
#define BESST_K_RADIX_HIST 1   /* k_radix_hist             (K3) */
#define BESST_K_RADIX_SCAN 2   /* k_radix_scan_hist        (K3) */
#define BESST_K_RADIX_SWEEP 3  /* k_radix_sweep, one per digit pass (K3) */
#define BESST_K_HEADS 4        /* k_head_count/k_scan_blocks/k_head_write (K4) */
#define BESST_K_EDGE_REDUCE 5  /* k_edge_reduce            (K4) */
#define BESST_K_EDGE_SCORE 6   /* k_ll_count/k_ll_write/k_score_keys: LL link space + sort keys (K5) */
#define BESST_K_FISHY 7        /* k_fishy_rekey */
#define BESST_K_METRICS 8      /* k_metrics_*              (K7) */
#define BESST_K_GAPEST 9       /* k_gapest_batch / k_edge_finalize (K6) */
#define BESST_K_TILE_SCAN 10   /* k_tile_reduce/k_chunk_resolve/k_tile_offsets (K2) */
#define BESST_K_COMPACT 11     /* k_compact_tuples         (K1) */
#define BESST_K_PARTITION 12   /* k_partition_count/scan/scatter (multi-GPU) */
#define BESST_K_KS_EVAL 13     /* k_ks_eval                (K5) */
#define BESST_K_KS_SORT 14     /* k_radix_sweep on the (edge, value) keys of K5, one per digit pass */
#define BESST_K_GROUP 15       /* k_group_blocks: block-local grouping of the tuple stream (K3') */
#define BESST_K_RUNS 16        /* k_run_count/k_scan_blocks64/k_run_write: heads and offsets over the sorted runs (K3') */
#define BESST_K_KS_BLOCK 17    /* k_ks_block: in-block sort + KS evaluation of the edges with <= 2048 links (K5') */
#define BESST_K_BAM_INFLATE 18 /* k_bgzf_inflate: one warp per BGZF block, raw deflate + CRC-32 (besst_bam_ingest) */
#define BESST_K_BAM_SCAN 19    /* k_bam_scan / k_bam_rescan: record boundaries per BGZF block */
#define BESST_K_BAM_DECODE 20  /* k_bam_decode: fixed core + CIGAR lengths -> record columns */
#define BESST_N_KERNEL_IDS 21
/* enabled: 0 off, 1 the launches of the LAST build, 2 accumulate over builds until besst_kernel_profile reads (and clears) them */
int besst_set_profiling(besst_ctx* ctx, int enabled);
int besst_kernel_profile(besst_ctx* ctx, int32_t* kernel_ids, float* ms, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* BESST_B200_H */
