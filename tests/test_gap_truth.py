"""Gap estimates against the SIMULATED TRUTH (SURVEY.md 8c iii): the only external anchor for the restated
mathstats GapEstimator (parity unpinned at that boundary).  The reference's test data and the synthetic
libraries encode every contig's genome interval in its name (`c<i>,pos:<start>-<end>,rc:<s>`), so the true
gap between two neighbouring contigs is next.start - this.end.  For every well-supported, scored edge between
two long scaffolds (the GapEstimator branch, CreateGraph.py:536-541) the ML gap must sit on the truth:

    |median error| <= 0.05 sigma,   median |error| <= 0.15 sigma,   and closer than the naive mu - mean(obs)

for both erf variants (A&S 7.1.26 = mathstats' own, and libm).  CPU: the C oracle on the reference's full BAMs
(where the reference tree exists) and on a synthetic library; GPU: the CUDA path on the synthetic library and
on the committed heads of the reference's BAMs."""
import io
import os
import re

import numpy as np
import pytest

import helpers
from besst_b200 import abi, synth
from besst_b200.contig_table import first_library_rows
from besst_b200.records import RecordBatch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_TESTDATA = "/root/reference/testdata"
VARIANTS = [abi.ERF_AS7126, abi.ERF_LIBM]


def _truth(names):
    a = np.zeros(len(names), np.int64)
    b = np.zeros(len(names), np.int64)
    for i, n in enumerate(names):
        m = re.search(r"pos:(\d+)-(\d+)", n)
        a[i], b[i] = int(m.group(1)), int(m.group(2))
    return a, b


def _gap_errors(res, rows, names, mu, min_links):
    """(ML gap - true gap, naive gap - true gap) over the scored edges between two long single-contig
    scaffolds that are neighbours in the truth and carry at least min_links links."""
    a, b = _truth(names)
    sel = ((res.flags & abi.EDGE_SCORED) != 0) & ((res.flags & abi.EDGE_BIG) != 0) & ((res.flags & abi.EDGE_NEGGAP) == 0) & (res.nr_links >= min_links)
    large = np.nonzero(rows["state"] == abi.CTG_LARGE)[0]   # first library: scaffold index i = i-th large contig
    cu, cv = large[res.edge_u[sel] >> 1], large[res.edge_v[sel] >> 1]
    true_gap = np.maximum(a[cv] - b[cu], a[cu] - b[cv])
    near = np.abs(true_gap) < mu
    ml = (res.gap[sel] - true_gap)[near]
    naive = (mu - res.obs_sum[sel] / res.nr_links[sel].astype(np.float64) - true_gap)[near]
    return ml, naive


def _assert_on_truth(ml, naive, sigma, label, min_edges):
    assert ml.size >= min_edges, "%s: only %d edges to judge" % (label, ml.size)
    med, mad = float(np.median(ml)), float(np.median(np.abs(ml)))
    assert abs(med) <= 0.05 * sigma, "%s: median gap error %.1f bp (sigma %.0f)" % (label, med, sigma)
    assert mad <= 0.15 * sigma, "%s: median |gap error| %.1f bp (sigma %.0f)" % (label, mad, sigma)
    assert abs(med) < abs(float(np.median(naive))), "%s: ML estimate no better than the naive one" % label


def _build(engine_build, batch, orientation, mu, sigma, read_len, threshold, contig_threshold, variant):
    rows, n_scaf, n_large = first_library_rows(np.asarray(batch.lengths), contig_threshold)
    params = abi.make_params(orientation, 11, read_len, mu, sigma, threshold, erf_variant=variant)
    return engine_build(rows, n_scaf, n_large, params, batch), rows


def _oracle_build(rows, n_scaf, n_large, params, batch):
    import oracle_lib
    return oracle_lib.graph_build(rows, n_scaf, params, batch)[0]


def _cuda_build(engine):
    def build(rows, n_scaf, n_large, params, batch):
        engine.set_contigs(rows, n_scaf, n_large)
        keep = []
        return engine.fetch(engine.build(params, abi.make_records(batch, keepalive=keep)))
    return build


def _synthetic():
    lib = synth.make_library(3000, 3_000_000, "rf", 3000.0, 500.0, 0.0, seed=4242)
    return lib.to_batch(), ("rf", 3000.0, 500.0, 100.0, 6000.0, 5000.0)


def _estimated_parameters(batch, engine):
    """what get_metrics leaves in param for a library run without -m/-s (testset2 has no Travis line)"""
    import contextlib
    from besst_b200 import libmetrics
    from besst_b200.records import BatchFile
    param = helpers.Param("/tmp", io.StringIO(), orientation="fr")
    with contextlib.redirect_stdout(io.StringIO()):
        libmetrics.get_metrics(BatchFile(batch), param, io.StringIO(), engine=engine)
    return ("fr", param.mean_ins_size, param.std_dev_ins_size, param.read_len, param.ins_size_threshold, param.contig_threshold)


def _check(build, batch, lib, variant, label, min_links, min_edges):
    orientation, mu, sigma, read_len, threshold, contig_threshold = lib
    res, rows = _build(build, batch, orientation, mu, sigma, read_len, threshold, contig_threshold, variant)
    ml, naive = _gap_errors(res, rows, batch.references, mu, min_links)
    _assert_on_truth(ml, naive, sigma, label, min_edges)


# ---- CPU: the C oracle (the restated estimator itself) ---------------------------------------------------------
@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_gaps_on_truth_synthetic(variant):
    batch, lib = _synthetic()
    _check(_oracle_build, batch, lib, variant, "synthetic/oracle", 20, 300)


@pytest.mark.skipif(not os.path.isdir(REF_TESTDATA), reason="needs the reference's testdata")
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("testset", ["testset1", "testset2"])
def test_oracle_gaps_on_truth_reference_testsets_full(testset, variant):
    from besst_b200 import bamio
    from oracle_engine import OracleEngine
    batch = bamio.read_bam_native(os.path.join(REF_TESTDATA, testset, "mapped.bam"))
    # testset1: the Travis command line (.travis.yml:14: -m 4000 -s 500 -k 3000 -T 6000); testset2: estimated
    lib = ("fr", 4000.0, 500.0, 100.0, 6000.0, 3000.0) if testset == "testset1" else _estimated_parameters(batch, OracleEngine())
    _check(_oracle_build, batch, lib, variant, testset + "/oracle", 20, 200)


# ---- GPU: the CUDA path through the C ABI -------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
def test_cuda_gaps_on_truth_synthetic(cuda_engine, variant):
    batch, lib = _synthetic()
    _check(_cuda_build(cuda_engine), batch, lib, variant, "synthetic/cuda", 20, 300)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("testset", ["testset1", "testset2"])
def test_cuda_gaps_on_truth_reference_testset_heads(cuda_engine, testset, variant):
    """the committed first 300 k / 250 k records of the reference's BAMs (the full files do not travel)"""
    batch = RecordBatch.load(os.path.join(GOLDEN, testset + "_head.npz"))
    lib = ("fr", 4000.0, 500.0, 100.0, 6000.0, 3000.0) if testset == "testset1" else ("fr", 2999.502067973291, 299.3939858319346, 100.0, 4795.865982964899, 4197.07801130103)
    _check(_cuda_build(cuda_engine), batch, lib, variant, testset + "/cuda", 10, 40)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("testset", ["testset1", "testset2"])
def test_oracle_gaps_on_truth_reference_testset_heads(testset, variant):
    batch = RecordBatch.load(os.path.join(GOLDEN, testset + "_head.npz"))
    lib = ("fr", 4000.0, 500.0, 100.0, 6000.0, 3000.0) if testset == "testset1" else ("fr", 2999.502067973291, 299.3939858319346, 100.0, 4795.865982964899, 4197.07801130103)
    _check(_oracle_build, batch, lib, variant, testset + "/oracle head", 10, 40)
