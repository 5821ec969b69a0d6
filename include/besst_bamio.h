/* besst_bamio.h -- C ABI of libbesst_bamio.so: sorted BAM file -> struct-of-arrays record columns.
 *
 * This is the ingest step in front of the hot path.  The reference walks the
 * BAM one pysam.AlignedRead at a time, three to four times per library (runBESST:162,
 * libmetrics.py:63,257,293, CreateGraph.py:111); here the file is inflated once by a pool of host
 * threads (BGZF blocks are independent raw-deflate streams) and the fixed-core fields the path needs
 * are decoded straight into the column layout of besst_records (include/besst_b200.h):
 *
 *   tid   refID          mtid  next_refID     pos   pos (0-based)     mpos  next_pos
 *   tlen  tlen           flag  flag           mapq  mapq
 *   qlen  l_seq minus leading/trailing soft clips  (pysam 0.8.4 `qlen`, CreateGraph.py:139)
 *   rlen  l_seq          alen  reference span of the CIGAR  (libmetrics.py:259-263, first 1000 records)
 *
 * Host-only library (g++, zlib, std::thread); no CUDA dependency.  The columns are owned by the
 * handle, 64-byte aligned, valid until besst_bam_close; besst_graph_build takes them as host
 * pointers (pass them through cudaHostRegister for full PCIe speed).
 */
#ifndef BESST_BAMIO_H
#define BESST_BAMIO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct besst_bam besst_bam;

typedef struct besst_bam_columns {
    int64_t n;            /* records */
    const int32_t* tid;
    const int32_t* mtid;
    const int32_t* pos;
    const int32_t* mpos;
    const int32_t* tlen;
    const int32_t* qlen;
    const uint16_t* flag;
    const uint8_t* mapq;
    const int32_t* rlen;  /* [min(n, head)] */
    const int32_t* alen;  /* [min(n, head)] */
    int64_t n_head;       /* records for which rlen / alen were kept */
    const uint32_t* packed; /* flag | mapq << 12 | qlen << 20 (besst_records.packed, include/besst_b200.h): what the graph
                               build uploads instead of flag / mapq / qlen.  NULL when some record does not fit
                               (flag >= 4096 or qlen >= 4096) */
} besst_bam_columns;

typedef struct besst_bam_stats {
    int64_t compressed_bytes, uncompressed_bytes, blocks, records;
    double seconds_inflate, seconds_decode, seconds_total;
    int32_t threads;
} besst_bam_stats;

int besst_bamio_abi_version(void);

/* Read the whole file: header + every record.  n_threads <= 0: hardware concurrency.
 * max_records < 0: no limit (else stop after that many records).  head_records: how many leading
 * records keep rlen/alen (the reference reads the first 1000).  NULL on failure: the message is in
 * err (if given). */
besst_bam* besst_bam_read(const char* path, int32_t n_threads, int64_t max_records, int64_t head_records,
                          char* err, int32_t err_len);

/* The same pass, streamed: every decoded window of records (a few hundred thousand) is handed to window_fn as
 * soon as it is ready and its column space is reused: memory stays bounded for a file of any size and a
 * consumer can move window k on (copy it to pinned staging and start its upload) before window k+1 is inflated.
 * window->n records starting at ordinal first_record; the pointers are valid only during the call (the call runs
 * on the caller's thread, between windows); b gives access to the header (besst_bam_n_refs ...) from the first
 * call on.  A non-zero return stops the pass: not an error, the handle is returned as usual and
 * besst_bam_stopped() says so (NULL always means a real failure, with the message in err).  The returned handle
 * holds the header, the statistics and rlen/alen of the first head_records records, but no record columns.
 * Every BGZF block's CRC32 is checked against the inflated bytes (BESST_BAMIO_NOCRC=1 skips the check). */
typedef int (*besst_bam_window_fn)(void* user, const besst_bam* b, const besst_bam_columns* window, int64_t first_record);
besst_bam* besst_bam_stream(const char* path, int32_t n_threads, int64_t max_records, int64_t head_records,
                            besst_bam_window_fn window_fn, void* user, char* err, int32_t err_len);

int64_t besst_bam_n_refs(const besst_bam* b);
const char* besst_bam_ref_name(const besst_bam* b, int64_t i);
int64_t besst_bam_ref_length(const besst_bam* b, int64_t i);
int besst_bam_get_columns(const besst_bam* b, besst_bam_columns* out);
int besst_bam_get_stats(const besst_bam* b, besst_bam_stats* out);
int besst_bam_stopped(const besst_bam* b);   /* 1: the window callback stopped the pass early */
void besst_bam_close(besst_bam* b);

#ifdef __cplusplus
}
#endif
#endif
