"""Shared test helpers: library construction, contig tables with multi-contig
scaffolds, and field-by-field comparison of two GraphResults."""
from __future__ import annotations

import numpy as np

from besst_b200 import abi
from besst_b200.contig_table import ContigTable
from besst_b200.objects import contig, scaffold

INT_FIELDS = ["edge_u", "edge_v", "nr_links", "obs_sum", "obs_sq", "first_idx", "row_ptr", "fishy", "flags",
              "obs_u", "obs_v", "aligned_len"]
FLOAT_FIELDS = ["score", "ks", "sd_obs", "sd_model"]
FLOAT_RTOL = 1e-6   # north-star tolerance for gap/score floats; integers are bit-exact


def first_library_objects(references, lengths, contig_threshold):
    """What InitializeObjects (CreateGraph.py:729-786) leaves behind."""
    Contigs, Scaffolds, small_contigs, small_scaffolds = {}, {}, {}, {}
    idx = 1
    for name, n in zip(references, lengths):
        c = contig(name, contig_direction=True, contig_position=0, contig_length=int(n))
        s = scaffold(idx, [c], int(n))
        c.scaffold = idx
        if n >= contig_threshold:
            Contigs[name], Scaffolds[idx] = c, s
        elif n > 0:
            small_contigs[name], small_scaffolds[idx] = c, s
        idx += 1
    return Contigs, Scaffolds, small_contigs, small_scaffolds


def later_library_objects(references, lengths, contig_threshold, seed, join=3, drop_every=17):
    """A state as MS.Algorithm would leave it for a later library: runs of
    consecutive contigs joined into multi-contig scaffolds with random
    directions and gaps, a few contigs removed (repeats)."""
    rng = np.random.default_rng(seed)
    Contigs, Scaffolds, small_contigs, small_scaffolds = {}, {}, {}, {}
    idx, i, n = 1, 0, len(references)
    while i < n:
        k = int(rng.integers(1, join + 1))
        members = []
        pos = 0
        for j in range(i, min(n, i + k)):
            if drop_every and j % drop_every == drop_every - 1:
                continue
            c = contig(references[j], contig_scaffold=idx, contig_direction=bool(rng.integers(0, 2)),
                       contig_position=pos, contig_length=int(lengths[j]))
            pos += int(lengths[j]) + int(rng.integers(1, 400))
            members.append(c)
        i += k
        if not members:
            continue
        s_len = members[-1].position + members[-1].length
        s = scaffold(idx, members, s_len)
        big = s_len >= contig_threshold
        for c in members:
            (Contigs if big else small_contigs)[c.name] = c
        (Scaffolds if big else small_scaffolds)[idx] = s
        idx += 1
    return Contigs, Scaffolds, small_contigs, small_scaffolds


def table_for(batch, objects):
    Contigs, Scaffolds, small_contigs, small_scaffolds = objects
    return ContigTable(batch.references, batch.lengths, Contigs, small_contigs, Scaffolds, small_scaffolds)


def assert_graph_equal(got, want, check_scores=True, label=""):
    assert got.n_edges == want.n_edges, "%s: edge count %d != %d" % (label, got.n_edges, want.n_edges)
    assert got.n_links == want.n_links, "%s: link count %d != %d" % (label, got.n_links, want.n_links)
    for f in INT_FIELDS:
        a, b = getattr(got, f), getattr(want, f)
        if f == "flags" and not check_scores:
            a, b = a & abi.EDGE_LL, b & abi.EDGE_LL
        assert np.array_equal(a, b), "%s: integer field %s differs at %s" % (label, f, np.nonzero(a != b)[0][:5])
    n_cnt = 12
    assert np.array_equal(got.counters[:n_cnt], want.counters[:n_cnt]), "%s: counters %s != %s" % (
        label, got.counters[:n_cnt], want.counters[:n_cnt])
    if check_scores:
        scored = (want.flags & abi.EDGE_SCORED) != 0
        assert np.array_equal(got.gap[scored], want.gap[scored]), "%s: gap differs" % label
        for f in FLOAT_FIELDS:
            a, b = getattr(got, f)[scored], getattr(want, f)[scored]
            assert np.array_equal(np.isnan(a), np.isnan(b)), "%s: NaN pattern of %s differs" % (label, f)
            ok = ~np.isnan(b)
            np.testing.assert_allclose(a[ok], b[ok], rtol=FLOAT_RTOL, atol=0, err_msg="%s: %s" % (label, f))


def max_rel_diff(a, b):
    ok = ~np.isnan(b)
    if not ok.any():
        return 0.0
    d = np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-300)
    return float(d.max())


# ---- canonical dump of a (G, G_prime) pair: the format of tests/golden/*.json.gz -----------------
def _crc(values):
    import zlib
    return zlib.crc32(np.asarray(list(values), dtype=np.int64).tobytes()) & 0xffffffff


def graph_signature(G):
    """Order-preserving, JSON-able dump of a scaffold graph as CreateGraph.PE
    returns it: node order, edge order (G.edges() iteration), every edge
    attribute; observation lists are folded to (length, crc32)."""
    nodes = [[n[0], n[1], G._node[n].get('length')] for n in G]
    edges = []
    for u, v, d in G.edges(data=True):
        row = {"u": [u[0], u[1]], "v": [v[0], v[1]], "nr_links": d.get('nr_links')}
        for key in ("obs", "obs_sq", "gap"):
            if key in d:
                row[key] = int(d[key])
        if "score" in d:
            row["score"] = float(d["score"])
        if "observations" in d:
            row["observations"] = [len(d["observations"]), _crc(d["observations"])]
        lists = {}
        for key, val in d.items():
            if not isinstance(key, str):   # per-scaffold lists are keyed by the int scaffold name
                lists[str(key)] = [len(val), _crc(val)]
        if lists:
            row["scaffold_lists"] = lists
        edges.append(row)
    return {"nodes": nodes, "edges": edges}


def assert_signature_equal(got, want, label="", score_rtol=FLOAT_RTOL):
    assert got["nodes"] == want["nodes"], "%s: node list/order differs" % label
    assert len(got["edges"]) == len(want["edges"]), "%s: %d edges != %d" % (label, len(got["edges"]), len(want["edges"]))
    for i, (a, b) in enumerate(zip(got["edges"], want["edges"])):
        sa, sb = a.pop("score", None), b.pop("score", None)
        assert a == b, "%s: edge %d differs:\n got %r\nwant %r" % (label, i, a, b)
        assert (sa is None) == (sb is None), "%s: edge %d score presence" % (label, i)
        if sa is not None:
            assert abs(sa - sb) <= score_rtol * abs(sb), "%s: edge %d score %r != %r" % (label, i, sa, sb)
            a["score"], b["score"] = sa, sb


PARAM_FLOAT_KEYS = ("read_len", "mean_ins_size", "std_dev_ins_size", "ins_size_threshold", "contig_threshold",
                    "mean_coverage", "std_dev_coverage", "contamination_ratio", "contamination_mean",
                    "contamination_stddev", "skewness", "skew_adj", "lognormal_mean", "lognormal_sigma")
PARAM_EXACT_KEYS = ("expected_links_over_mean_plus_stddev", "edgesupport", "scaffold_indexer", "tot_assembly_length",
                    "current_N50", "current_L50", "lognormal")


def param_signature(param):
    out = {}
    for k in PARAM_FLOAT_KEYS + PARAM_EXACT_KEYS:
        v = getattr(param, k, None)
        if isinstance(v, (np.floating, np.integer)):
            v = v.item()
        out[k] = v
    return out


def assert_param_equal(got, want, label="", rtol=1e-9):
    for k in PARAM_EXACT_KEYS:
        assert got.get(k) == want.get(k), "%s: param.%s %r != %r" % (label, k, got.get(k), want.get(k))
    for k in PARAM_FLOAT_KEYS:
        a, b = got.get(k), want.get(k)
        if b is None or b is False or a is None or a is False:
            assert a == b or (not a and not b), "%s: param.%s %r != %r" % (label, k, a, b)
        else:
            assert abs(a - b) <= rtol * max(abs(b), 1e-300), "%s: param.%s %r != %r" % (label, k, a, b)


def object_signature(Contigs, Scaffolds, small_contigs, small_scaffolds):
    def crows(d):
        return [[name, c.scaffold, bool(c.direction), int(c.position), int(c.length), None if c.coverage is None else float(c.coverage)] for name, c in d.items()]

    def srows(d):
        return [[name, [c.name for c in s.contigs], int(s.s_length)] for name, s in d.items()]
    return {"Contigs": crows(Contigs), "small_contigs": crows(small_contigs),
            "Scaffolds": srows(Scaffolds), "small_scaffolds": srows(small_scaffolds)}


class Param(object):
    """Minimal stand-in for BESST.Parameter.parameter (Parameter.py:24-105): a bag
    of attributes initialised the way runBESST:88-158 does for one library."""

    def __init__(self, outdir, information, **opts):
        o = dict(orientation="fr", mean=None, stddev=None, threshold=None, minsize=None, readlen=None,
                 edgesupport=None, min_mapq=11, covcutoff=None, lower_covcutoff=0.001, duplicate=True,
                 extendpaths=True, no_score=False)
        o.update(opts)
        self.scaffold_indexer = 1
        self.no_score = o["no_score"]
        self.min_mapq = o["min_mapq"]
        self.max_contig_overlap = 200
        self.cov_cutoff = o["covcutoff"]
        self.lower_cov_cutoff = o["lower_covcutoff"]
        self.plots = False
        self.development = False
        self.print_scores = False
        self.first_lib = True
        self.pass_number = 1
        self.bamfile = "in_memory.bam"
        self.orientation = o["orientation"]
        self.mean_ins_size = o["mean"]
        self.std_dev_ins_size = o["stddev"]
        self.ins_size_threshold = o["threshold"]
        self.contig_threshold = o["minsize"]
        self.edgesupport = o["edgesupport"]
        self.read_len = o["readlen"]
        self.output_directory = outdir
        self.information_file = information
        self.detect_haplotype = False
        self.hapl_ratio = 1.3
        self.hapl_threshold = 3
        self.detect_duplicate = o["duplicate"]
        self.extend_paths = o["extendpaths"]
        self.lognormal = False
        self.contamination_ratio = None
        self.tot_assembly_length = None


class FakeSeq(object):
    """Sequence stand-in (only len() and slicing are used on this path)."""
    __slots__ = ("n",)

    def __init__(self, n):
        self.n = int(n)

    def __len__(self):
        return self.n

    def __getitem__(self, s):
        if isinstance(s, slice):
            lo, hi, _ = s.indices(self.n)
            return "N" * max(0, hi - lo)
        return "N"


def run_dropin(batch, opts, engine, state=None, run_libmetrics=True, fasta_lengths=None, bam_path=None, param_overrides=None):
    """One library pass through the drop-in entry points (besst_b200.libmetrics.get_metrics
    + besst_b200.CreateGraph.PE) with `engine` behind them.  Mirrors
    oracle/ref_harness.run_reference so that the two outputs compare field by field."""
    import io
    import tempfile
    from besst_b200 import CreateGraph as CG, libmetrics
    from besst_b200.records import BatchFile
    outdir = tempfile.mkdtemp(prefix="besst_b200_")
    info = io.StringIO()
    param = Param(outdir, info, **opts)
    param.first_lib = state is None
    for k, v in (param_overrides or {}).items():
        setattr(param, k, v)
    if fasta_lengths is None:
        fasta_lengths = dict(zip(batch.references, batch.lengths))
    C_dict = {name: FakeSeq(n) for name, n in fasta_lengths.items()}
    if state is None:
        Contigs, Scaffolds, small_contigs, small_scaffolds = {}, {}, {}, {}
    else:
        Contigs, Scaffolds = state["Contigs"], state["Scaffolds"]
        small_contigs, small_scaffolds = state["small_contigs"], state["small_scaffolds"]
        param.scaffold_indexer = state["scaffold_indexer"]
        param.tot_assembly_length = state["tot_assembly_length"]
    bam_file = BatchFile(batch) if bam_path is None else bam_path   # a path goes through the native ingest library
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        if run_libmetrics:
            libmetrics.get_metrics(bam_file, param, info, engine=engine)
        G, G_prime = CG.PE(Contigs, Scaffolds, info, C_dict, param, small_contigs, small_scaffolds, bam_file, engine=engine)
    return {"G": graph_signature(G), "G_prime": graph_signature(G_prime), "param": param_signature(param),
            "objects": object_signature(Contigs, Scaffolds, small_contigs, small_scaffolds),
            "information": info.getvalue()}


def state_for_later_library(batch, contig_threshold, seed):
    Contigs, Scaffolds, small_contigs, small_scaffolds = later_library_objects(
        batch.references, batch.lengths, contig_threshold, seed=seed)
    n = max(list(Scaffolds) + list(small_scaffolds)) + 1
    return dict(Contigs=Contigs, Scaffolds=Scaffolds, small_contigs=small_contigs, small_scaffolds=small_scaffolds,
                scaffold_indexer=n, tot_assembly_length=int(sum(batch.lengths)))


def state_snapshot(Contigs, Scaffolds, small_contigs, small_scaffolds, param):
    """JSON-able picture of the four object dicts (dict ORDER kept: it defines the scaffold numbering and the node
    insertion order) as a scaffolding pass leaves them for the next library (runBESST:199-231)."""
    def crows(d):
        return [[name, c.scaffold, bool(c.direction), int(c.position), int(c.length)] for name, c in d.items()]

    def srows(d):
        return [[name, [c.name for c in s.contigs], int(s.s_length)] for name, s in d.items()]
    return {"Contigs": crows(Contigs), "small_contigs": crows(small_contigs), "Scaffolds": srows(Scaffolds),
            "small_scaffolds": srows(small_scaffolds), "scaffold_indexer": int(param.scaffold_indexer),
            "tot_assembly_length": int(param.tot_assembly_length)}


def state_from_snapshot(snap):
    """-> the `state` argument of run_dropin / ref_harness.run_reference (fresh objects, BESST's own classes when importable)"""
    from besst_b200.objects import classes
    contig_cls, scaffold_cls = classes()
    objs = {}
    out = {"Contigs": {}, "small_contigs": {}, "Scaffolds": {}, "small_scaffolds": {}}
    for key in ("Contigs", "small_contigs"):
        for name, scaf, direction, position, length in snap[key]:
            c = contig_cls(name, contig_scaffold=scaf, contig_direction=direction, contig_position=position, contig_length=length)
            c.sequence = FakeSeq(length)
            objs[name] = c
            out[key][name] = c
    for key in ("Scaffolds", "small_scaffolds"):
        for name, members, s_length in snap[key]:
            out[key][name] = scaffold_cls(name, [objs[m] for m in members], s_length)
    out["scaffold_indexer"] = snap["scaffold_indexer"]
    out["tot_assembly_length"] = snap["tot_assembly_length"]
    return out
