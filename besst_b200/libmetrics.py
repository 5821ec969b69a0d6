"""Drop-in for `BESST.libmetrics.get_metrics` (reference libmetrics.py:226-432).

Same signature, same fields set on `param`, same lines written to
`Information`.  The two capped BAM scans -- insert-size sampling (:283-304) and
the contamination scan (:49-84), both filtered by the record predicates of
bam_parser.py:22-29 -- run as CUDA kernels that rank qualifying records in BAM
order, cut at the reference's 1e6-sample caps and build integer histograms
(besst_libmetrics in include/besst_b200.h); the trim loop (:316-332), skewness
(:340-341) and getdistr (:141-223) are then evaluated over histogram bins
behind the same ABI call.  What stays here is the host bookkeeping: read
length from the first 1000 records (:246-273), the thresholds (:275-281,
:404-409), the lognormal switch (:360-390), the contamination verdict
(:112-128) and the LIBRARY STATISTICS block (:415-431).

There is no CPU fallback: without the CUDA library `get_metrics` raises.
"""
from __future__ import annotations

import math
import sys

import numpy as np

from . import abi
from .contig_table import largest_reference_mask
from .records import as_batch, as_file

ISIZE_DICT_CAP = 1 << 22   # bins of param.empirical_distribution fetched from the engine


def _set_thresholds(param):
    """-T and -k from mean/sd (libmetrics.py:275-281 and :404-409)."""
    param.ins_size_threshold = param.mean_ins_size + 6 * param.std_dev_ins_size
    if param.extend_paths:
        param.contig_threshold = param.mean_ins_size + 4 * param.std_dev_ins_size
    else:
        param.contig_threshold = param.mean_ins_size + (param.std_dev_ins_size / float(param.mean_ins_size)) * param.std_dev_ins_size


def estimate_read_length(batch):
    """Mean of rlen (alen when rlen == 0) over the first 1000 records, as a float
    (libmetrics.py:246-266).  None when the file holds fewer than 1000 records."""
    if len(batch) < 1000:
        return None
    rlen = batch.rlen if batch.rlen is not None else batch.qlen
    alen = batch.alen if batch.alen is not None else batch.qlen
    r = np.asarray(rlen[:1000], dtype=np.int64)
    a = np.asarray(alen[:1000], dtype=np.int64)
    return int(np.where(r != 0, r, a).sum()) / float(1000)


def _chunk_modes(adj):
    """mode of the chunk-summed distribution for the 21 window sizes
    (libmetrics.py:203-209), for the Information lines only."""
    out = []
    for chunk in range(1, 102, 5):
        sums = np.add.reduceat(adj, np.arange(0, adj.shape[0], chunk)) if adj.shape[0] else np.zeros(1)
        out.append((chunk, (int(np.argmax(sums)) + 0.5) * chunk))
    return out


def metric_rows(lengths):
    """Contig rows carrying only what the metrics kernels gather: in_largest,
    the membership of a tid in the 1000 longest references (:231-233)."""
    rows = np.zeros(len(lengths), dtype=abi.CONTIG_ROW_DTYPE)
    rows["in_largest"] = largest_reference_mask(lengths)
    rows["length"] = np.asarray(lengths, dtype=np.int64)
    return rows


def _get_metrics_whole_file(path, param, Information, engine):
    """the library's head does not lie inside rank 0's part: read the whole file on host threads for this call"""
    from .bamio import read_bam_native
    from .records import BatchFile
    return get_metrics(BatchFile(read_bam_native(path)), param, Information, engine=engine)


def get_metrics(bam_file, param, Information, engine=None):
    bam_file = as_file(bam_file, engine)   # a path: decoded once by the native ingest library
    cont_names = bam_file.references
    cont_lengths = [int(x) for x in bam_file.lengths]
    param.lognormal = False
    try:
        bam_file.fetch(cont_names[0])
    except ValueError:
        sys.stderr.write('Need indexed bamfiles, index file should be located in the same directory as the BAM file\nterminating..\n')
        sys.exit(0)
    batch = as_batch(bam_file)
    parts = getattr(batch, "dist_info", None)   # a process group ingested the file in parts (dist.ingest_bam_distributed)
    if parts is not None:
        # the metrics read a BAM-order PREFIX of the library (the first 1000 records, then the capped sampling scans): when
        # rank 0's part covers that prefix its result is the whole file's and is broadcast; otherwise every rank falls back
        # to the host reader for this call (PE keeps working on the parts)
        from .dist import host_group, rank0_value
        total = sum(parts["counts"])
        if min(total, 1000) > parts["counts"][0]:
            return _get_metrics_whole_file(parts["path"], param, Information, engine)

    if not param.read_len:
        read_len = estimate_read_length(batch) if parts is None else rank0_value(lambda: estimate_read_length(batch), parts["rank"], host_group())
        if parts is not None and read_len is None and total >= 1000:
            return _get_metrics_whole_file(parts["path"], param, Information, engine)
        if read_len is None:
            sys.stderr.write('Did not get sufficient readmappings to calculate\
             read_length from mappings. Got {0} mappings. Please provide this parameter or more importantly\
             check why almost no reads are mapping to the contigs.\nterminating..\n'.format(len(batch)))
            sys.exit(0)
        param.read_len = read_len

    if param.mean_ins_size and param.std_dev_ins_size and not param.ins_size_threshold:
        _set_thresholds(param)
        print('-T', param.ins_size_threshold, '-t', param.contig_threshold, file=Information)

    want_isize = not param.mean_ins_size
    if engine is None:
        from .engine import default_engine
        engine = default_engine()
    params = abi.make_params(param.orientation, param.min_mapq, param.read_len, param.mean_ins_size,
                             param.std_dev_ins_size, param.ins_size_threshold,
                             detect_duplicate=param.detect_duplicate, extend_paths=param.extend_paths,
                             no_score=param.no_score)
    if parts is None:
        rc, m, adj = engine.libmetrics(metric_rows(cont_lengths), params, batch, cont_lengths, want_isize,
                                       cap=ISIZE_DICT_CAP)
    else:
        def on_rank0():
            rc0, m0, adj0 = engine.libmetrics(metric_rows(cont_lengths), params, batch, cont_lengths, want_isize, cap=ISIZE_DICT_CAP)
            # complete: the capped scans stopped inside rank 0's part (or it is the whole file)
            complete = int(m0.records_scanned) < len(batch) or total == len(batch)
            return complete, rc0, {f: getattr(m0, f) for f, _ in abi.LibMetricsOut._fields_}, np.array(adj0)
        complete, rc, fields, adj = rank0_value(on_rank0, parts["rank"], host_group())
        if not complete:
            return _get_metrics_whole_file(parts["path"], param, Information, engine)
        import types
        m = types.SimpleNamespace(**fields)

    if want_isize:
        line = "Estimating insert size from {0} mappings with quality over --min_mapq {1}.".format(m.n_samples + 1, param.min_mapq)
        print(line)
        print(line, file=Information)
        if rc == 1:   # libmetrics.py:311-314
            sys.stderr.write('To few valid read alignments exists to compute mean and variance of library (need at least 1000 observations). Got only ' + str(m.n_samples) + ' valid alignments. Please specify -m and -s to the program. \nPrinting out scaffolds produced in earlier steps...')
            sys.stderr.write('\nterminating...\n')
            sys.exit(0)
        print('Mean before filtering :', m.mean_before, file=Information)
        print('Std_est  before filtering: ', m.sd_before, file=Information)
        print('Mean converged:', m.mean_converged, file=Information)
        print('Std_est converged: ', m.sd_converged, file=Information)
        param.mean_ins_size = m.mean_converged
        param.std_dev_ins_size = m.sd_converged
        param.skewness = m.skewness
        print('Skewness of distribution: ', param.skewness, file=Information)
        for chunk, mode in _chunk_modes(adj):
            print("mode for chunk size ", chunk, " : ", mode, file=Information)
        print("Choosing mode:", m.mode_adj)
        print('mu_adjusted:{0}, sigma_adjusted:{1}, skewness_adjusted:{2}'.format(m.mu_adj, m.sigma_adj, m.skew_adj))
        param.skew_adj = m.skew_adj
        param.empirical_distribution = dict(zip(range(adj.shape[0]), adj.tolist()))
        print('Mean of getdistr adjusted distribution: ', m.mu_adj, file=Information)
        print('Sigma of getdistr adjusted distribution: ', m.sigma_adj, file=Information)
        print('Skewness of getdistr adjusted distribution: ', m.skew_adj, file=Information)
        print('Median of getdistr adjusted distribution: ', m.median_adj, file=Information)
        print('Mode of getdistr adjusted distribution: ', m.mode_adj, file=Information)
        print('Using mean and stddev of getdistr adjusted distribution from here: ', m.mu_adj, m.sigma_adj, file=Information)
        param.mean_ins_size = m.mu_adj
        param.std_dev_ins_size = m.sigma_adj
        if param.skew_adj > 0.5 and math.log(m.median_adj) > math.log(m.mode_adj):   # :360-390
            print('Mode on getdistr adjusted: ', m.mode_adj, file=Information)
            print("Median on getdistr adjusted:", m.median_adj, file=Information)
            print("mode adj:", m.mode_adj)
            print("median adj", m.median_adj)
            param.lognormal_mean = math.log(m.median_adj)
            param.lognormal_sigma = math.sqrt(param.lognormal_mean - math.log(m.mode_adj))
            print('Lognormal mean getdistr adjusted: ', param.lognormal_mean, file=Information)
            print("Lognormal stddev getdistr adjusted", param.lognormal_sigma, file=Information)
            param.lognormal = True

    if not param.ins_size_threshold:
        _set_thresholds(param)

    # contamination verdict (libmetrics.py:86-128)
    n_contamine = float(m.cont_n)
    if m.cont_n_before > 2:   # :91-110
        print('Contamine mean before filtering :', m.cont_mean_before, file=Information)
        print('Contamine stddev before filtering: ', m.cont_sd_before, file=Information)
        print('Contamine mean converged:', m.cont_mean, file=Information)
        print('Contamine std_est converged: ', m.cont_sd, file=Information)
    ratio = 2 * n_contamine / float(m.cont_mapped) if m.cont_mapped > 0 else 0
    if m.cont_mean >= param.mean_ins_size or m.cont_sd >= param.std_dev_ins_size or ratio <= 0.05:
        param.contamination_ratio = False
        param.contamination_mean = 0
        param.contamination_stddev = 0
    else:
        param.contamination_mean = m.cont_mean
        param.contamination_stddev = m.cont_sd
        param.contamination_ratio = ratio
    if hasattr(bam_file, "reset"):
        bam_file.reset()

    print('', file=Information)
    print('LIBRARY STATISTICS', file=Information)
    print('Mean of library set to:', param.mean_ins_size, file=Information)
    print('Standard deviation of library set to: ', param.std_dev_ins_size, file=Information)
    print('MP library PE contamination:', file=Information)
    print('Contamine rate (rev comp oriented) estimated to: ', param.contamination_ratio, file=Information)
    print('lib contamine mean (avg fragmentation size): ', param.contamination_mean, file=Information)
    print('lib contamine stddev: ', param.contamination_stddev, file=Information)
    print('Number of contamined reads used for this calculation: ', n_contamine, file=Information)
    print('-T (library insert size threshold) set to: ', param.ins_size_threshold, file=Information)
    print('-k set to (Scaffolding with contigs larger than): ', param.contig_threshold, file=Information)
    print('Number of links required to create an edge: ', param.edgesupport, file=Information)
    print('Maximum identical contig-end overlap-length to merge of contigs that are adjacent in a scaffold: ', param.max_contig_overlap, file=Information)
    print('Read length set to: ', param.read_len, file=Information)
    print('', file=Information)
    return ()
