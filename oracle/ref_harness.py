"""ORACLE / TEST INFRASTRUCTURE (tier 1) -- run the reference's OWN bytecode.

Imports /root/reference/BESST/{CreateGraph,libmetrics,bam_parser,e_nr_links}.py
unmodified, after placing stand-ins for the three dependencies that are absent
from this image into sys.modules:

  pysam      -> oracle/pysam_stub.py (pre-decoded records)
  networkx   -> oracle/nx1compat.py  (1.x API over the installed 3.x)
  mathstats  -> oracle/mathstats_restated/ (restated, PARITY UNPINNED)

and mirrors the per-library prologue of runBESST:88-182 (parameter object,
contig_index, get_metrics, PE).  Integers produced this way (edges, link
counts, counters, observation lists) are authoritative; floats that pass
through `mathstats` are "vs our restatement".

Only runs where /root/reference exists (this container); it produces the
golden fixtures under tests/golden/ via oracle/make_golden.py.  Never imported
by the product path.
"""
from __future__ import annotations

import io
import os
import re
import sys
import tempfile
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("BESST_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "BESST", "CreateGraph.py"))


_loaded = None


def load_reference():
    """Import the reference modules with the shims installed; returns a
    namespace (CG, libmetrics, Parameter, Contig, Scaffold, e_nr_links, ...)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    ms = os.path.join(HERE, "mathstats_restated")
    if ms not in sys.path:
        sys.path.insert(0, ms)
    import nx1compat
    import pysam_stub
    nx1compat.install()
    pysam_stub.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import BESST.CreateGraph as CG
    import BESST.libmetrics as libmetrics
    import BESST.Parameter as Parameter
    import BESST.Contig as Contig
    import BESST.Scaffold as Scaffold
    import BESST.e_nr_links as e_nr_links
    import BESST.bam_parser as bam_parser
    _loaded = SimpleNamespace(CG=CG, libmetrics=libmetrics, Parameter=Parameter, Contig=Contig,
                              Scaffold=Scaffold, e_nr_links=e_nr_links, bam_parser=bam_parser,
                              pysam=pysam_stub, nx=sys.modules["networkx"])
    return _loaded


class FakeSeq(object):
    """Contig sequence stand-in: only len() and slicing are used on the hot
    path (CreateGraph.py:737,758; GenerateOutput.py:41-79)."""
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


DEFAULT_OPTIONS = dict(orientation="fr", mean=None, stddev=None, threshold=None, minsize=None,
                       readlen=None, edgesupport=None, min_mapq=11, covcutoff=None,
                       lower_covcutoff=0.001, duplicate=True, extendpaths=True, no_score=False,
                       haplratio=1.3, haplthreshold=3)


def make_param(ref, opts, outdir, first_lib=True, pass_number=1):
    """runBESST:88-158 for one library."""
    o = dict(DEFAULT_OPTIONS)
    o.update(opts)
    param = ref.Parameter.parameter()
    param.scaffold_indexer = 1
    param.multiprocess = False
    param.no_score = o["no_score"]
    param.score_cutoff = 1.5
    param.max_extensions = None
    param.NO_ILP = False
    param.FASTER_ILP = False
    param.dfs_traversal = True
    param.print_scores = False
    param.min_mapq = o["min_mapq"]
    param.max_contig_overlap = 200
    param.cov_cutoff = o["covcutoff"]
    param.lower_cov_cutoff = o["lower_covcutoff"]
    param.development = False
    param.plots = False
    param.first_lib = first_lib
    param.path_threshold = 100000
    param.pass_number = pass_number
    param.bamfile = "in_memory.bam"
    param.orientation = o["orientation"]
    param.mean_ins_size = o["mean"]
    param.ins_size_threshold = o["threshold"]
    param.edgesupport = o["edgesupport"]
    param.read_len = o["readlen"]
    param.output_directory = outdir
    param.std_dev_ins_size = o["stddev"]
    param.contig_threshold = o["minsize"]
    param.hapl_ratio = o["haplratio"]
    param.hapl_threshold = o["haplthreshold"]
    param.detect_haplotype = False
    param.detect_duplicate = o["duplicate"]
    param.extend_paths = o["extendpaths"]
    return param


def graph_dump(G):
    """Canonical, order-preserving dump of a networkx graph."""
    nodes = [(n, dict(G._node[n])) for n in list(G)]
    edges = []
    for u, v in list(G.edges()):
        edges.append((u, v, {k: (list(val) if isinstance(val, list) else val) for k, val in G[u][v].items()}))
    return {"nodes": nodes, "edges": edges}


_COUNTER_PATTERNS = {
    "fishy": r"NR OF FISHY READ LINKS:\s+(\d+)",
    "count": r"Number of USEFUL READS \(reads mapping to different contigs uniquly\):\s+(\d+)",
    "non_unique": r"that maps to different contigs \(filtered out from scaffolding\):\s+(\d+)",
    "too_long": r"Reads with too large insert size from \"USEFUL READS\" \(filtered out\):\s+(\d+)",
    "duplicates": r"Number of duplicated reads indicated and removed:\s+(\d+)",
    "initial_edges_G": r"Initial number of edges in G \(the graph with large contigs\):\s+(\d+)",
    "initial_edges_G_prime": r"Initial number of edges in G_prime \(the full graph of all contigs before removal of repats\):\s+(\d+)",
    "bug_edges_removed": r"Number of BWA buggy edges removed:\s+(\d+)",
    "low_support_removed_G": r"Removed (\d+) edges from graph G of border contigs",
    "high_density_removed": r"Removed total of (\d+) edges in high density areas",
    "low_support_removed_G_prime": r"Removed an additional of (\d+) edges with low support",
}


def run_reference(batch, opts, fasta_lengths=None, state=None, run_libmetrics=True):
    """One library pass of the reference: get_metrics + PE.

    batch          RecordBatch (records + header)
    opts           runBESST options for this library (see DEFAULT_OPTIONS)
    fasta_lengths  {contig name: length}; default = the BAM header
    state          None for a first library; otherwise a dict with prebuilt
                   Contigs/Scaffolds/small_contigs/small_scaffolds (reference
                   objects) and scaffold_indexer, as MS.Algorithm would leave
                   them for the next library
    Returns a dict (see bottom of this function)."""
    ref = load_reference()
    outdir = tempfile.mkdtemp(prefix="besst_ref_")
    info = io.StringIO()
    param = make_param(ref, opts, outdir, first_lib=state is None)
    param.information_file = info
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

    bam_file = ref.pysam.Samfile(batch)
    param.contig_index = dict(zip(range(len(bam_file.references)), bam_file.references))

    captured = {}
    orig_counters = ref.CG.counters
    orig_bug = ref.CG.RemoveBugEdges

    class CapturingCounters(orig_counters):
        def __init__(self, *a, **k):
            orig_counters.__init__(self, *a, **k)
            captured["counter"] = self

    def capturing_bug(G, G_prime, fishy_edges, param_, Information):
        captured["fishy_edges"] = dict(fishy_edges)
        captured["G_pre"] = graph_dump(G)
        captured["G_prime_pre"] = graph_dump(G_prime)
        return orig_bug(G, G_prime, fishy_edges, param_, Information)

    ref.CG.counters = CapturingCounters
    ref.CG.RemoveBugEdges = capturing_bug
    stdout = sys.stdout
    sys.stdout = io.StringIO()
    try:
        if run_libmetrics:
            ref.libmetrics.get_metrics(bam_file, param, info)
        metrics_param = {k: getattr(param, k, None) for k in (
            "read_len", "mean_ins_size", "std_dev_ins_size", "ins_size_threshold", "contig_threshold",
            "skewness", "skew_adj", "lognormal", "lognormal_mean", "lognormal_sigma",
            "contamination_ratio", "contamination_mean", "contamination_stddev")}
        empirical = getattr(param, "empirical_distribution", None)
        G, G_prime = ref.CG.PE(Contigs, Scaffolds, info, C_dict, param, small_contigs, small_scaffolds, bam_file)
    finally:
        ref.CG.counters = orig_counters
        ref.CG.RemoveBugEdges = orig_bug
        libmetrics_stdout = sys.stdout.getvalue()
        sys.stdout = stdout

    text = info.getvalue()
    counters = {}
    for key, pat in _COUNTER_PATTERNS.items():
        m = re.search(pat, text)
        counters[key] = int(m.group(1)) if m else None
    c = captured.get("counter")
    if c is not None:
        counters.update(count=c.count, non_unique=c.non_unique, non_unique_for_scaf=c.non_unique_for_scaf,
                        duplicates_raw=c.nr_of_duplicates, too_long=c.reads_with_too_long_insert)

    def contig_rows(d):
        return [(name, o.scaffold, bool(o.direction), int(o.position), int(o.length), o.coverage) for name, o in d.items()]

    def scaffold_rows(d):
        return [(name, [c_.name for c_ in s.contigs], int(s.s_length)) for name, s in d.items()]

    return {
        "param": {k: getattr(param, k, None) for k in (
            "read_len", "mean_ins_size", "std_dev_ins_size", "ins_size_threshold", "contig_threshold",
            "mean_coverage", "std_dev_coverage", "expected_links_over_mean_plus_stddev", "edgesupport",
            "contamination_ratio", "contamination_mean", "contamination_stddev", "scaffold_indexer",
            "tot_assembly_length", "current_N50", "current_L50", "lognormal", "no_score", "extend_paths",
            "orientation", "min_mapq", "detect_duplicate")},
        "metrics_param": metrics_param,
        "empirical_distribution": empirical,
        "counters": counters,
        "G": graph_dump(G), "G_prime": graph_dump(G_prime),
        "G_pre": captured.get("G_pre"), "G_prime_pre": captured.get("G_prime_pre"),
        "fishy_edges": captured.get("fishy_edges"),
        "Contigs": contig_rows(Contigs), "small_contigs": contig_rows(small_contigs),
        "Scaffolds": scaffold_rows(Scaffolds), "small_scaffolds": scaffold_rows(small_scaffolds),
        "information": text, "stdout": libmetrics_stdout, "outdir": outdir,
        "objects": dict(Contigs=Contigs, Scaffolds=Scaffolds, small_contigs=small_contigs,
                        small_scaffolds=small_scaffolds, param=param, G=G, G_prime=G_prime),
    }
