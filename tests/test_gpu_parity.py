"""-m gpu: the CUDA path against the C oracle, through the C ABI, on seeded
synthetic libraries (bit-exact integers, 1e-6 relative floats)."""
import numpy as np
import pytest

import helpers
import oracle_lib
from besst_b200 import abi, synth

pytestmark = pytest.mark.gpu

CASES = [
    # config, objects, params overrides
    ("tiny", "first", {}),
    ("small_pe", "first", {}),
    ("small_mp", "first", {}),
    ("small_mp_cont", "first", {}),
    ("small_pe", "later", {}),
    ("small_mp", "later", {"detect_duplicate": False}),
    ("small_pe", "first", {"no_score": True}),
    ("small_mp", "first", {"extend_paths": False}),
    ("small_mp", "later", {"no_score": True, "extend_paths": False, "min_mapq": 0}),
    ("small_pe", "later", {"read_len": 99.37}),
]


def _params(lib, over):
    mu, sd = lib.mu, lib.sigma
    kw = dict(orientation=lib.orientation, min_mapq=11, read_len=100.0, mean_ins_size=mu, std_dev_ins_size=sd,
              ins_size_threshold=mu + 6 * sd)
    kw.update(over)
    return abi.make_params(**kw), mu + 4 * sd


@pytest.mark.parametrize("config,objects,over", CASES)
def test_graph_build_matches_oracle(cuda_engine, config, objects, over):
    lib = synth.make_config(config)
    batch = lib.to_batch()
    params, contig_threshold = _params(lib, over)
    if objects == "first":
        objs = helpers.first_library_objects(batch.references, batch.lengths, contig_threshold)
    else:
        objs = helpers.later_library_objects(batch.references, batch.lengths, contig_threshold, seed=7)
    table = helpers.table_for(batch, objs)
    want, tuples, fishy, consistent = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
    assert consistent
    got = cuda_engine.graph_build(table, params, batch)
    helpers.assert_graph_equal(got, want, label="%s/%s/%s" % (config, objects, over))
    assert want.n_edges > 0 and want.n_links > 0


@pytest.mark.parametrize("over", [
    {"ins_size_threshold": 5.0e6},     # 23 value bits: KS keys do not fit the in-block path, 64-bit device-wide sorts
    {"ins_size_threshold": 70000.0},   # 17 value bits
    {"ins_size_threshold": 26.0},      # nothing passes obs1 > 25 and obs2 > 25 and obs1 + obs2 < 26: no links at all
    {"ins_size_threshold": float("nan")},
    {"min_mapq": 61},                  # no record passes the link filter
])
def test_threshold_extremes(cuda_engine, over):
    lib = synth.make_config("small_mp")
    batch = lib.to_batch()
    params, contig_threshold = _params(lib, over)
    objs = helpers.first_library_objects(batch.references, batch.lengths, contig_threshold)
    table = helpers.table_for(batch, objs)
    want, _, _, consistent = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
    assert consistent
    got = cuda_engine.graph_build(table, params, batch)
    helpers.assert_graph_equal(got, want, label=str(over))


def test_library_without_inter_contig_pairs(cuda_engine):
    lib, batch, params, table = _setup("small_pe")
    from besst_b200.records import RecordBatch
    same = RecordBatch(tid=batch.tid, mtid=batch.tid.copy(), pos=batch.pos, mpos=batch.mpos, tlen=batch.tlen, qlen=batch.qlen,
                       flag=batch.flag, mapq=batch.mapq, references=batch.references, lengths=batch.lengths)
    want, _, _, _ = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, same)
    got = cuda_engine.graph_build(table, params, same)
    helpers.assert_graph_equal(got, want, label="no inter-contig pairs")
    assert got.n_edges == 0 and got.counters[abi.CNT_VALID] == len(batch)


def _setup(config, over=None, objects="first"):
    lib = synth.make_config(config)
    batch = lib.to_batch()
    params, contig_threshold = _params(lib, over or {})
    if objects == "first":
        objs = helpers.first_library_objects(batch.references, batch.lengths, contig_threshold)
    else:
        objs = helpers.later_library_objects(batch.references, batch.lengths, contig_threshold, seed=7)
    return lib, batch, params, helpers.table_for(batch, objs)


@pytest.mark.parametrize("n", [0, 1, 2, 31, 1023, 1024, 1025, 4097])
def test_ragged_and_empty_inputs(cuda_engine, n):
    lib, batch, params, table = _setup("small_mp")
    sub = batch.slice(50000, 50000 + n)
    want, _, _, _ = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, sub)
    got = cuda_engine.graph_build(table, params, sub)
    helpers.assert_graph_equal(got, want, label="n=%d" % n)


def test_records_with_absent_or_negative_contigs(cuda_engine):
    lib, batch, params, table = _setup("small_pe", objects="later")   # later: some contigs removed
    rng = np.random.default_rng(1)
    tid = batch.tid.copy()
    mtid = batch.mtid.copy()
    tid[rng.random(len(batch)) < 0.01] = -1
    mtid[rng.random(len(batch)) < 0.01] = -1
    mtid[rng.random(len(batch)) < 0.001] = len(batch.references) + 5    # out of range: dropped like tid < 0
    from besst_b200.records import RecordBatch
    bad = RecordBatch(tid=tid, mtid=mtid, pos=batch.pos, mpos=batch.mpos, tlen=batch.tlen, qlen=batch.qlen,
                      flag=batch.flag, mapq=batch.mapq, references=batch.references, lengths=batch.lengths)
    want, _, _, _ = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, bad)
    got = cuda_engine.graph_build(table, params, bad)
    helpers.assert_graph_equal(got, want, label="absent")


def test_edges_with_more_links_than_one_warp_sorts(cuda_engine):
    """Few long contigs -> edges with thousands of links (the CTA-per-edge path of K5)."""
    lib = synth.make_library(12, 400000, "rf", 3000.0, 500.0, seed=99)
    batch = lib.to_batch()
    params = abi.make_params("rf", 11, 100.0, 3000.0, 500.0, 6000.0)
    objs = helpers.first_library_objects(batch.references, batch.lengths, 1000.0)
    table = helpers.table_for(batch, objs)
    want, _, _, _ = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
    assert want.nr_links.max() > 4096
    got = cuda_engine.graph_build(table, params, batch)
    helpers.assert_graph_equal(got, want, label="big edges")


@pytest.mark.parametrize("config", ["small_mp", "small_mp_cont", "mixed_edge_sizes"])
def test_ks_block_path_equals_global_sort_path(cuda_engine, config, monkeypatch):
    """KS statistic: edges with <= 2048 links are sorted and evaluated inside one CTA, larger ones by
    two device-wide key sorts; BESST_KS=global sends everything through the latter."""
    if config == "mixed_edge_sizes":   # few contigs of very different lengths: edges from a handful to > 2048 links
        lib = synth.make_library(40, 300000, "rf", 3000.0, 500.0, seed=123)
        batch = lib.to_batch()
        params = abi.make_params("rf", 11, 100.0, 3000.0, 500.0, 6000.0)
        objs = helpers.first_library_objects(batch.references, batch.lengths, 1000.0)
        table = helpers.table_for(batch, objs)
    else:
        lib, batch, params, table = _setup(config)
    want, _, _, _ = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
    monkeypatch.setenv("BESST_KS", "global")
    got = cuda_engine.graph_build(table, params, batch)
    helpers.assert_graph_equal(got, want, label="global KS %s" % config)
    monkeypatch.delenv("BESST_KS")
    got = cuda_engine.graph_build(table, params, batch)
    helpers.assert_graph_equal(got, want, label="block KS %s" % config)
    if config == "mixed_edge_sizes":
        ll = (want.flags & abi.EDGE_LL) != 0
        assert (want.nr_links[ll] > 2048).any() and (want.nr_links[ll] <= 2048).any()


def test_device_resident_records_equal_host_records(cuda_engine):
    import torch
    lib, batch, params, table = _setup("small_mp_cont")
    host = cuda_engine.graph_build(table, params, batch)
    cols = {k: torch.from_numpy(np.ascontiguousarray(v).view(np.int16) if k == "flag" else np.ascontiguousarray(v)).cuda()
            for k, v in batch.device_arrays().items()}
    ptrs = {k: v.data_ptr() for k, v in cols.items()}
    ptrs["n"] = len(batch)
    torch.cuda.synchronize()
    cuda_engine.set_table(table)
    dev = cuda_engine.fetch(cuda_engine.build(params, abi.make_records(ptrs, on_device=True)))
    helpers.assert_graph_equal(dev, host, label="device-resident")


@pytest.mark.parametrize("slice_records", [128, 1024, 128 * 77])
def test_sliced_host_copy_equals_one_pass(cuda_engine, slice_records, monkeypatch):
    """Host-buffer calls copy the record columns slice by slice and run the record kernel on each slice
    as it lands (H2D overlap): the result must not depend on the slice size."""
    from besst_b200.engine import CudaEngine
    lib, batch, params, table = _setup("small_mp_cont")
    want = cuda_engine.graph_build(table, params, batch)
    monkeypatch.setenv("BESST_SLICE_RECORDS", str(slice_records))
    eng = CudaEngine()
    try:
        for n in (len(batch), 5 * slice_records + 1, slice_records + 127):
            sub = batch.slice(0, min(n, len(batch)))
            ref = cuda_engine.graph_build(table, params, sub) if n < len(batch) else want
            got = eng.graph_build(table, params, sub)
            helpers.assert_graph_equal(got, ref, label="slice=%d n=%d" % (slice_records, n))
    finally:
        eng.close()


def test_fetch_view_equals_fetch(cuda_engine):
    lib, batch, params, table = _setup("small_pe")
    cuda_engine.set_table(table)
    keep = []
    sizes = cuda_engine.build(params, abi.make_records(batch, keepalive=keep))
    a = cuda_engine.fetch(sizes)
    b = cuda_engine.fetch_view(sizes)
    helpers.assert_graph_equal(b, a, label="view")
    # views are reused by the next view call, copies are not
    sub = batch.slice(0, 2000)
    keep2 = []
    s2 = cuda_engine.build(params, abi.make_records(sub, keepalive=keep2))
    c = cuda_engine.fetch_view(s2)
    assert c.n_edges <= a.n_edges and c.counters[abi.CNT_VALID] <= a.counters[abi.CNT_VALID]
    helpers.assert_graph_equal(cuda_engine.fetch(s2), c, label="view after rebuild")


@pytest.mark.parametrize("config,objects", [("small_mp", "first"), ("small_pe", "later"), ("small_mp_cont", "first")])
def test_radix_bucket_fallback_equals_run_merge_bucket(cuda_engine, config, objects, monkeypatch):
    """The edge bucket has two implementations: block-local grouping + merge of the run descriptors
    (default) and a device-wide radix sort (fallback for input without local order)."""
    lib, batch, params, table = _setup(config, objects=objects)
    want, _, _, _ = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
    monkeypatch.setenv("BESST_BUCKET", "radix")
    got = cuda_engine.graph_build(table, params, batch)
    helpers.assert_graph_equal(got, want, label="radix bucket %s" % config)
    monkeypatch.delenv("BESST_BUCKET")
    got = cuda_engine.graph_build(table, params, batch)
    helpers.assert_graph_equal(got, want, label="run-merge bucket %s" % config)


@pytest.mark.parametrize("order", ["reversed", "window", "shuffled"])
def test_bucket_on_tuple_streams_without_bam_order(cuda_engine, order, monkeypatch):
    """links_to_graph accepts any tuple order (multi-GPU exchange).  A fully shuffled stream has more
    than 512 edges per block: the run-merge bucket must notice and hand over to the radix bucket."""
    import torch
    lib, batch, params, table = _setup("small_mp")
    cuda_engine.set_table(table)
    keep = []
    cuda_engine.links_extract(params, abi.make_records(batch, keepalive=keep))
    tuples = cuda_engine.links_tuples_host()
    rng = np.random.default_rng(5)
    n = tuples.shape[0]
    if order == "reversed":
        perm = np.arange(n)[::-1].copy()
    elif order == "window":
        perm = np.arange(n)
        for a in range(0, n, 300):
            rng.shuffle(perm[a:a + 300])
    else:
        perm = rng.permutation(n)
    t = torch.from_numpy(np.ascontiguousarray(tuples[perm]).view(np.int32).reshape(-1, 4)).cuda()
    res = {}
    for mode in ("radix", "runs"):
        if mode == "radix":
            monkeypatch.setenv("BESST_BUCKET", "radix")
        else:
            monkeypatch.delenv("BESST_BUCKET", raising=False)
        sizes = cuda_engine.links_to_graph(params, t.data_ptr(), n, None, 0)
        res[mode] = cuda_engine.fetch(sizes)
    helpers.assert_graph_equal(res["runs"], res["radix"], label="tuple order %s" % order)
    # same edges and link multisets as the BAM-ordered build
    sizes = cuda_engine.links_to_graph(params, torch.from_numpy(tuples.view(np.int32).reshape(-1, 4)).cuda().data_ptr(), n, None, 0)
    ref = cuda_engine.fetch(sizes)
    assert np.array_equal(ref.edge_u, res["runs"].edge_u) and np.array_equal(ref.nr_links, res["runs"].nr_links)
    assert np.array_equal(ref.obs_sum, res["runs"].obs_sum) and np.array_equal(ref.obs_sq, res["runs"].obs_sq)


def test_bam_file_to_graph_through_native_ingest(cuda_engine, tmp_path):
    """BAM file -> libbesst_bamio.so (threaded inflate + decode) -> host columns -> graph build: equals the
    oracle on the in-memory records the file was written from."""
    from besst_b200 import bamio
    from test_bamio import write_bam
    lib, batch, params, table = _setup("tiny")
    recs = []
    for i in range(len(batch)):
        q = int(batch.qlen[i])
        recs.append(dict(tid=int(batch.tid[i]), pos=int(batch.pos[i]), mapq=int(batch.mapq[i]), flag=int(batch.flag[i]), l_seq=q,
                         mtid=int(batch.mtid[i]), mpos=int(batch.mpos[i]), tlen=int(batch.tlen[i]), cigar=[(0, q)] if q else []))
    path = str(tmp_path / "lib.bam")
    write_bam(path, list(zip(batch.references, [int(x) for x in batch.lengths])), recs, block_bytes=20000)
    nat = bamio.read_bam_native(path, threads=4)
    assert len(nat) == len(batch) and nat.references == list(batch.references)
    want, _, _, _ = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
    got = cuda_engine.graph_build(table, params, nat)
    helpers.assert_graph_equal(got, want, label="bam ingest")


def test_two_slices_with_halo_equal_one_pass(cuda_engine):
    """The multi-GPU decomposition on one GPU: records cut into BAM-order slices, each
    extracted with the previous slice's last CreateEdge observation as halo, tuples
    concatenated in slice order == the single-pass tuple stream."""
    lib, batch, params, table = _setup("small_mp")
    cuda_engine.set_table(table)
    keep = []
    n = len(batch)
    whole = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
    tuples_all, counters_all, aligned_all = [], np.zeros(abi.N_COUNTERS, np.int64), np.zeros(len(batch.references), np.int64)
    halo = (-1, -1)
    for lo, hi in ((0, n // 3), (n // 3, n // 3 + 7), (n // 3 + 7, n)):
        p = abi.make_params(lib.orientation, 11, 100.0, lib.mu, lib.sigma, lib.mu + 6 * lib.sigma, halo=halo)
        nt = cuda_engine.links_extract(p, abi.make_records(batch.slice(lo, hi), keepalive=keep))
        tuples_all.append(cuda_engine.links_tuples_host())
        assert tuples_all[-1].shape[0] == nt
        aligned, counters = cuda_engine.links_partials()
        halo = (int(counters[abi.CNT_LAST_OBS1]), int(counters[abi.CNT_LAST_OBS2]))
        counters_all[:8] += counters[:8]
        aligned_all += aligned
    got = np.concatenate(tuples_all)
    assert np.array_equal(got, whole[1])
    assert np.array_equal(counters_all[:8], whole[0].counters[:8])
    assert np.array_equal(aligned_all, whole[0].aligned_len)
    assert halo == (int(whole[0].counters[abi.CNT_LAST_OBS1]), int(whole[0].counters[abi.CNT_LAST_OBS2]))


@pytest.mark.parametrize("erf_variant", [abi.ERF_AS7126, abi.ERF_LIBM])
def test_gapest_batch_matches_oracle(cuda_engine, erf_variant):
    rng = np.random.default_rng(11)
    n = 20000
    for mean, sd, r in ((3000.0, 500.0, 100.0), (550.0, 50.0, 99.37), (8000.0, 1200.0, 150.0)):
        p = abi.make_params("fr", 11, r, mean, sd, mean + 6 * sd, erf_variant=erf_variant)
        mo = rng.uniform(2 * r, mean + 3 * sd, n)
        l1 = rng.integers(int(2 * sd) + 1, 60000, n).astype(np.int32)
        l2 = rng.integers(int(2 * sd) + 1, 60000, n).astype(np.int32)
        gap, sd_out = cuda_engine.gapest_batch(p, mo, l1, l2)
        gap_o, sd_o = oracle_lib.gapest_batch(p, mo, l1, l2)
        if erf_variant == abi.ERF_AS7126:
            assert np.array_equal(gap, gap_o)
        else:   # device erf() and glibc erf() may differ by an ulp: a bisection step can flip by 1 bp
            assert np.abs(gap - gap_o).max() <= 1 and (gap != gap_o).mean() < 1e-3
        same = gap == gap_o
        if erf_variant == abi.ERF_AS7126:
            np.testing.assert_allclose(sd_out[same], sd_o[same], rtol=helpers.FLOAT_RTOL, atol=0)
        else:   # erf differences cancel in g(d): an ulp of erf can show up as 1e-4 of the sd in the far tails
            rel = np.abs(sd_out[same] - sd_o[same]) / np.maximum(np.abs(sd_o[same]), 1e-300)
            assert (rel > helpers.FLOAT_RTOL).mean() < 1e-3 and rel.max() < 1e-2


def test_libmetrics_matches_oracle_and_hits_the_sample_cap(cuda_engine):
    """> 1e6 qualifying read2 records on the 1000 longest contigs: both capped scans
    (libmetrics.py:283-304 and :49-84) must cut at the same BAM-order prefix."""
    from besst_b200 import libmetrics
    lib = synth.make_library(50, 1400000, "fr", 550.0, 50.0, seed=321)
    batch = lib.to_batch()
    params = abi.make_params("fr", 11, 100.0, 0.0, 0.0, 0.0)
    rows = libmetrics.metric_rows(batch.lengths)
    rc_o, m_o, adj_o = oracle_lib.libmetrics(rows, params, batch, batch.lengths, True)
    rc, m, adj = cuda_engine.libmetrics(rows, params, batch, batch.lengths, True)
    assert rc == rc_o == 0
    assert m_o.n_samples == 1000000
    for f in ("n_samples", "n_trimmed", "median_adj", "mode_adj", "n_bins", "cont_mapped", "cont_n", "records_scanned", "cont_n_before"):
        assert getattr(m, f) == getattr(m_o, f), f
    for f in ("mean_before", "sd_before", "mean_converged", "sd_converged", "skewness", "mu_adj", "sigma_adj",
              "skew_adj", "cont_mean", "cont_sd", "cont_mean_before", "cont_sd_before"):
        a, b = getattr(m, f), getattr(m_o, f)
        assert abs(a - b) <= 1e-9 * max(abs(b), 1e-12), (f, a, b)
    np.testing.assert_allclose(adj, adj_o, rtol=1e-12, atol=0)


def test_too_few_insert_size_samples_is_reported_not_fatal(cuda_engine):
    from besst_b200 import libmetrics
    lib, batch, _, _ = _setup("tiny")
    sub = batch.slice(0, 2000)
    params = abi.make_params("fr", 11, 100.0, 0.0, 0.0, 0.0)
    rows = libmetrics.metric_rows(sub.lengths)
    rc, m, _ = cuda_engine.libmetrics(rows, params, sub, sub.lengths, True)
    rc_o, m_o, _ = oracle_lib.libmetrics(rows, params, sub, sub.lengths, True)
    assert (rc, m.n_samples) == (rc_o, m_o.n_samples) and rc == 1 and m.n_samples <= 1000


def test_full_size_config2_matches_oracle(cuda_engine):
    """BASELINE.json configs[1] at full size (10k contigs / 20 M PE pairs): the whole CSR
    against the sequential C oracle."""
    import torch
    from besst_b200.contig_table import first_library_rows
    lib = synth.make_config("config2", device="cuda", with_names=False)
    batch = lib.to_batch()
    rows, n_scaf, n_large = first_library_rows(lib.lengths.numpy(), lib.mu + 4 * lib.sigma)
    params = abi.make_params("fr", 11, 100.0, lib.mu, lib.sigma, lib.mu + 6 * lib.sigma)
    want, _, _, consistent = oracle_lib.graph_build(rows, n_scaf, params, batch)
    assert consistent
    cuda_engine.set_contigs(rows, n_scaf, n_large)
    ptrs = {k: v.data_ptr() for k, v in lib.cols.items()}
    ptrs["n"] = lib.n_records
    torch.cuda.synchronize()   # the generator ran on torch's stream, the engine has its own: device pointers must be ready
    got = cuda_engine.fetch(cuda_engine.build(params, abi.make_records(ptrs, on_device=True)))
    helpers.assert_graph_equal(got, want, label="config2 full")
    del lib
    torch.cuda.empty_cache()


def test_full_size_config3_against_oracle(cuda_engine):
    """BASELINE.json configs[2] at FULL size (100k contigs / 202 M MP pairs, 404 M records, the bench workload):
    every integer of the CSR, every observation list, counters and coverages bit-exact against the sequential
    C oracle (~11 s on one core), gaps equal, scores to 1e-6."""
    import gc
    import torch
    from besst_b200.contig_table import first_library_rows
    lib = synth.make_config("config3", device="cuda", with_names=False)
    rows, n_scaf, n_large = first_library_rows(lib.lengths.numpy(), lib.mu + 4 * lib.sigma)
    params = abi.make_params("rf", 11, 100.0, lib.mu, lib.sigma, lib.mu + 6 * lib.sigma)
    cuda_engine.set_contigs(rows, n_scaf, n_large)
    ptrs = {k: v.data_ptr() for k, v in lib.cols.items()}
    ptrs["n"] = lib.n_records
    torch.cuda.synchronize()   # the generator ran on torch's stream, the engine has its own: device pointers must be ready
    got = cuda_engine.fetch(cuda_engine.build(params, abi.make_records(ptrs, on_device=True)))
    assert lib.n_records > 400_000_000 and got.n_links > 50_000_000
    batch = lib.to_batch()
    del lib
    torch.cuda.empty_cache()
    want, _, _, consistent = oracle_lib.graph_build(rows, n_scaf, params, batch)
    assert consistent
    helpers.assert_graph_equal(got, want, label="config3 full")
    del batch, want, got
    gc.collect()


def test_full_size_config3_invariants(cuda_engine):
    """BASELINE.json configs[2] (100k contigs / 200 M MP pairs) at half scale: size-independent
    properties of the CSR -- sorted unique edge keys, row_ptr consistent with nr_links,
    per-edge sums equal to the payload, counters balance, and idempotence."""
    import torch
    from besst_b200.contig_table import first_library_rows
    lib = synth.make_config("config3", device="cuda", with_names=False, scale=0.5)
    rows, n_scaf, n_large = first_library_rows(lib.lengths.numpy(), lib.mu + 4 * lib.sigma)
    params = abi.make_params("rf", 11, 100.0, lib.mu, lib.sigma, lib.mu + 6 * lib.sigma)
    cuda_engine.set_contigs(rows, n_scaf, n_large)
    ptrs = {k: v.data_ptr() for k, v in lib.cols.items()}
    ptrs["n"] = lib.n_records
    torch.cuda.synchronize()   # the generator ran on torch's stream, the engine has its own: device pointers must be ready
    rec = abi.make_records(ptrs, on_device=True)
    a = cuda_engine.fetch(cuda_engine.build(params, rec))
    b = cuda_engine.fetch(cuda_engine.build(params, rec))
    for f in helpers.INT_FIELDS + ["gap"]:
        assert np.array_equal(getattr(a, f), getattr(b, f)), "not idempotent: %s" % f
    assert np.array_equal(a.score, b.score, equal_nan=True)
    key = (a.edge_u.astype(np.int64) << 32) | a.edge_v
    assert (np.diff(key) > 0).all() and (a.edge_u < a.edge_v).all()
    assert a.row_ptr[0] == 0 and a.row_ptr[-1] == a.n_links and np.array_equal(np.diff(a.row_ptr), a.nr_links)
    tot = a.obs_u.astype(np.int64) + a.obs_v
    assert np.array_equal(np.add.reduceat(tot, a.row_ptr[:-1]), a.obs_sum)
    assert np.array_equal(np.add.reduceat(tot * tot, a.row_ptr[:-1]), a.obs_sq)
    assert (a.obs_u > 25).all() and (a.obs_v > 25).all() and (tot < lib.mu + 6 * lib.sigma).all()
    c = a.counters
    ll_links = int(a.nr_links[(a.flags & abi.EDGE_LL) != 0].sum())
    assert c[abi.CNT_COUNT] == a.n_links + ll_links          # LL links are counted twice (CreateGraph.py:176,183)
    assert len(set(a.first_idx.tolist())) == a.n_edges and a.first_idx.max() < a.n_links
    scored = (a.flags & abi.EDGE_SCORED) != 0
    assert np.array_equal(scored, (a.flags & abi.EDGE_LL) != 0)
    s = a.score[scored]
    assert ((s == 0) | ((s > 1.0) & (s <= 2.0))).all()
    # true gaps of the generator are in [0, 1500]: the typical ML estimate on a well-supported edge is in that range
    well = scored & (a.nr_links >= 50) & ((a.flags & abi.EDGE_NEGGAP) == 0)
    assert well.sum() > 1000 and 0 < np.median(a.gap[well]) < 1500
    del lib
    torch.cuda.empty_cache()


def edge_dest_numpy(u, v, world):
    """besst_links.cu edge_dest: murmur3 finaliser of (u << 32 | v) mod world."""
    with np.errstate(over="ignore"):
        x = (u.astype(np.uint64) << np.uint64(32)) | v.astype(np.uint64)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
    return (x % np.uint64(world)).astype(np.int64)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_by_edge_hash_is_stable(cuda_engine, world):
    import torch
    lib, batch, params, table = _setup("small_mp_cont")
    cuda_engine.set_table(table)
    keep = []
    n = cuda_engine.links_extract(params, abi.make_records(batch, keepalive=keep))
    tuples = cuda_engine.links_tuples_host()
    fishy = cuda_engine.links_fishy_host()
    out_t = torch.zeros(max(n, 1) * 4, dtype=torch.int32, device="cuda")
    out_f = torch.zeros(max(len(fishy), 1), dtype=torch.int64, device="cuda")
    out_o = torch.zeros(max(n, 1), dtype=torch.int32, device="cuda")
    tc, fc = cuda_engine.links_partition(world, out_t.data_ptr(), out_f.data_ptr(), out_o.data_ptr())
    assert tc.sum() == n and fc.sum() == len(fishy)
    got = out_t.cpu().numpy().view(abi.LINK_TUPLE_DTYPE)[:n]
    dest = edge_dest_numpy(tuples["u"], tuples["v"], world)
    order = np.argsort(dest, kind="stable")
    assert np.array_equal(np.bincount(dest, minlength=world), tc)
    assert np.array_equal(got, tuples[order])
    assert np.array_equal(out_o.cpu().numpy()[:n], order)
    got_f = out_f.cpu().numpy().view(np.uint64)[:len(fishy)]
    fdest = edge_dest_numpy((fishy >> np.uint64(32)).astype(np.uint32), (fishy & np.uint64(0xffffffff)).astype(np.uint32), world)
    assert np.array_equal(np.bincount(fdest, minlength=world), fc)
    # fishy keys are appended with atomics (order-free): compare bucket contents as multisets
    start = 0
    for d in range(world):
        assert np.array_equal(np.sort(got_f[start:start + fc[d]]), np.sort(fishy[fdest == d]))
        start += fc[d]


def test_trsk_sd_batch_matches_oracle(cuda_engine):
    rng = np.random.default_rng(2)
    n = 5000
    p = abi.make_params("rf", 11, 100.0, 3000.0, 500.0, 6000.0)
    gap = rng.integers(-800, 3500, n).astype(np.float64)
    l1 = rng.integers(1001, 50000, n).astype(np.int32)
    l2 = rng.integers(1001, 50000, n).astype(np.int32)
    got = cuda_engine.trsk_sd_batch(p, gap, l1, l2)
    L = oracle_lib.lib()
    want = np.array([L.besst_oracle_tr_sk_std_dev(3000.0, 500.0, 100.0, float(a), float(b), float(g), abi.ERF_AS7126)
                     for a, b, g in zip(l1, l2, gap)])
    np.testing.assert_allclose(got, want, rtol=helpers.FLOAT_RTOL, atol=0)


def test_gapest_batch_with_fractional_contig_lengths(cuda_engine):
    """mathstats takes float lengths (c1 = c2 = mean + 4*stdDev at MakeScaffolds.py:68): fp64 across the ABI."""
    rng = np.random.default_rng(5)
    n = 2000
    p = abi.make_params("rf", 11, 99.37, 2987.4142135, 512.7182818, 6000.0)
    mo = rng.uniform(200.0, 3400.0, n)
    l1 = rng.uniform(1100.0, 40000.0, n)
    l2 = rng.uniform(1100.0, 40000.0, n)
    gap, sd = cuda_engine.gapest_batch(p, mo, l1, l2)
    gap_o, sd_o = oracle_lib.gapest_batch(p, mo, l1, l2)
    assert np.array_equal(gap, gap_o)
    np.testing.assert_allclose(sd, sd_o, rtol=helpers.FLOAT_RTOL, atol=0)
    L = oracle_lib.lib()
    want = np.array([L.besst_oracle_gap_estimator(2987.4142135, 512.7182818, 99.37, float(m), float(a), float(b), abi.ERF_AS7126)
                     for m, a, b in zip(mo[:200], l1[:200], l2[:200])])
    assert np.array_equal(gap[:200], want)   # the lengths were not truncated on the way


def test_scalar_dropins(cuda_engine):
    from besst_b200 import param_est
    L = oracle_lib.lib()
    assert param_est.GapEstimator(3000.0, 500.0, 100.0, 2500.0, 6000, 7000, engine=cuda_engine) == \
        L.besst_oracle_gap_estimator(3000.0, 500.0, 100.0, 2500.0, 6000.0, 7000.0, abi.ERF_AS7126)
    sd = param_est.tr_sk_std_dev(3000.0, 500.0, 100.0, 6000, 7000, 420, engine=cuda_engine)
    assert sd == pytest.approx(L.besst_oracle_tr_sk_std_dev(3000.0, 500.0, 100.0, 6000.0, 7000.0, 420.0, abi.ERF_AS7126), rel=1e-6)


@pytest.mark.parametrize("mean,sd,r", [(3000.0, 500.0, 100.0), (550.0, 50.0, 100.0), (8000.0, 1200.0, 150.0), (350.0, 100.0, 75.0),
                                       # what MakeScaffolds.py:68 really passes: get_metrics' estimates (mu_adj, sigma_adj, mean read length)
                                       (2987.4142135, 512.7182818, 99.37), (561.0307, 47.93, 100.0)])
def test_precalc_table_of_long_contig_gaps_equals_restated_mathstats(cuda_engine, mean, sd, r):
    """GC.PreCalcMLvaluesOfdLongContigs (MakeScaffolds.py:68): every d of the table in one kernel launch."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "mathstats_restated"))
    from mathstats.normaldist.truncatedskewed import param_est as ref
    from besst_b200 import param_est
    want = ref.PreCalcMLvaluesOfdLongContigs(mean, sd, r)
    got = param_est.PreCalcMLvaluesOfdLongContigs(mean, sd, r, engine=cuda_engine)
    assert got == want and len(want) > 100
    assert list(got.keys()) == list(want.keys())   # same insertion order
    ds = np.array([-200.0, 0.0, 17.5, 900.0])
    f = param_est.func_of_d_batch(mean, sd, r, ds, 6000, 4100, engine=cuda_engine)
    for d, v in zip(ds, f):
        assert v == pytest.approx(ref.funcDGeneral(float(d), mean, sd, 6000, 4100, r)[0], rel=1e-12)


def test_distributed_nccl_equals_oracle():
    """Needs >= 2 GPUs on the box (gpurun --gpus 2): the real NCCL all-to-all path."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU on this box: the N>1 host logic is covered by tests/test_dist_gloo.py")
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    world = min(torch.cuda.device_count(), 4)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(here, "dist_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_lognormal_gapest_batch_matches_oracle(cuda_engine):
    """besst_gapest_lognormal_batch (one warp per edge over the raw observations) against the C restatement:
    ragged edges from 1 to 3000 observations, short and long contigs, negative and large gaps."""
    rng = np.random.default_rng(12)
    mu, sigma, r = 8.0, 0.25, 100.0
    samples, row_ptr, l1, l2 = [], [0], [], []
    for i in range(300):
        n = int(rng.choice([1, 2, 5, 31, 32, 33, 100, 700, 3000]))
        d = int(rng.integers(-150, 3000))
        s = np.clip(np.rint(rng.lognormal(mu, sigma, n) - d), 210, None).astype(np.int64)
        samples += s.tolist()
        row_ptr.append(len(samples))
        l1.append(float(rng.integers(1200, 40000)))
        l2.append(float(rng.integers(1200, 40000)) + 0.5 * (i % 2))
    got = cuda_engine.gapest_lognormal_batch(mu, sigma, r, samples, row_ptr, l1, l2)
    want = oracle_lib.gapest_lognormal_batch(mu, sigma, r, samples, row_ptr, l1, l2)
    assert np.array_equal(got, want), np.nonzero(got != want)[0][:10]


def test_lognormal_scoring_branch_cuda_equals_oracle(cuda_engine):
    """CreateGraph.lognormal_rescore (CreateGraph.py:485-493,523-531) with the CUDA engine against the oracle engine"""
    import math
    from oracle_engine import OracleEngine
    from besst_b200 import CreateGraph as CG
    lib, batch, params, table = _setup("small_mp")

    class P(object):
        mean_ins_size, read_len = lib.mu, 100.0
        lognormal_sigma = 0.17
        lognormal_mean = math.log(lib.mu) - 0.17 ** 2 / 2
        empirical_distribution = {x: math.exp(-((x - lib.mu) / lib.sigma) ** 2 / 2) for x in range(200, 6001)}
    got = cuda_engine.graph_build(table, params, batch)
    want = OracleEngine().graph_build(table, params, batch)
    CG.lognormal_rescore(got, table, P, cuda_engine)
    CG.lognormal_rescore(want, table, P, OracleEngine())
    helpers.assert_graph_equal(got, want, label="lognormal")
    assert (got.gap[(got.flags & abi.EDGE_BIG) != 0] != 0).any()


@pytest.mark.parametrize("config", ["small_mp_cont", "small_pe"])
def test_packed_record_column_equals_plain_columns(cuda_engine, config):
    """besst_records.packed (flag | mapq << 12 | qlen << 20, 20 B/record) through K1's TMA path, its ragged leftover
    path, the sliced host upload and device-resident pointers: identical to the three plain columns and to the oracle."""
    import torch
    from besst_b200.records import RecordBatch
    lib, batch, params, table = _setup(config)
    rng = np.random.default_rng(3)
    batch.qlen = rng.integers(30, 151, len(batch)).astype(np.int32)   # ragged read lengths: the coverage sums depend on them
    want = cuda_engine.graph_build(table, params, batch)
    ref, _, _, _ = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
    helpers.assert_graph_equal(want, ref, label="plain")
    packed = RecordBatch(references=batch.references, lengths=batch.lengths, **batch.device_arrays()).with_packed()
    assert packed.packed is not None and packed.packed.dtype == np.uint32
    got = cuda_engine.graph_build(table, params, packed)
    helpers.assert_graph_equal(got, want, label="packed host")
    # only the packed column and the four coordinates are touched: poison the plain ones
    poisoned = RecordBatch(references=batch.references, lengths=batch.lengths, packed=packed.packed, **batch.device_arrays())
    poisoned.flag = np.zeros_like(batch.flag); poisoned.mapq = np.zeros_like(batch.mapq); poisoned.qlen = np.zeros_like(batch.qlen)
    helpers.assert_graph_equal(cuda_engine.graph_build(table, params, poisoned), want, label="packed only")
    # device-resident, unaligned start (leftover path) and a ragged tail
    cols = {k: torch.from_numpy(np.ascontiguousarray(getattr(packed, k)[3:-5])).cuda() for k in ("tid", "mtid", "pos", "mpos")}
    cols["packed"] = torch.from_numpy(packed.packed[3:-5].view(np.int32).copy()).cuda()
    ptrs = {k: v.data_ptr() for k, v in cols.items()}
    ptrs["n"] = len(batch) - 8
    torch.cuda.synchronize()
    cuda_engine.set_table(table)
    dev = cuda_engine.fetch(cuda_engine.build(params, abi.make_records(ptrs, on_device=True)))
    sub = batch.slice(3, len(batch) - 5)
    helpers.assert_graph_equal(dev, cuda_engine.graph_build(table, params, sub), label="packed device")
    assert abi.pack_record_columns(np.array([4096]), np.array([0]), np.array([100])) is None
    assert abi.pack_record_columns(np.array([99]), np.array([60]), np.array([5000])) is None


def test_resident_contig_tables_switch_without_reupload(cuda_engine):
    """besst_contigs_select: the libraries of a run keep their contig tables side by side in HBM (runBESST:143-231:
    library k+1 sees the scaffolds built from library k); switching back and forth gives the results of fresh uploads."""
    lib, batch, params, first = _setup("small_mp", objects="first")
    _, _, _, later = _setup("small_mp", objects="later")
    want_first = cuda_engine.graph_build(first, params, batch)
    want_later = cuda_engine.graph_build(later, params, batch)
    assert want_first.n_edges != want_later.n_edges
    keep = []
    rec = abi.make_records(batch, keepalive=keep)
    try:
        cuda_engine.select_table(1); cuda_engine.set_table(first)
        cuda_engine.select_table(2); cuda_engine.set_table(later)
        for slot, want in ((1, want_first), (2, want_later), (1, want_first), (2, want_later)):
            cuda_engine.select_table(slot)
            helpers.assert_graph_equal(cuda_engine.fetch(cuda_engine.build(params, rec)), want, label="slot %d" % slot)
        with pytest.raises(Exception):
            cuda_engine.select_table(8)
    finally:
        cuda_engine.select_table(0)
