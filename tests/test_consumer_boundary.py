"""Consumer-level proof of the drop-in boundary: the graphs returned by
besst_b200.CreateGraph.PE are handed to the reference's OWN consumer,
BESST.MakeScaffolds.Algorithm (MakeScaffolds.py:49-130, called at runBESST:199), under the
networkx-1.x API the reference requires, and must lead to the very scaffolds the reference's own
CreateGraph.PE output leads to -- for one library and for a two-library sequence
(runBESST:143-231: PE -> Algorithm -> next library's PE with first_lib = False, where CleanObjects
:788-810 and the general PosDir cases :1031-1048 see real multi-contig scaffolds).

CPU only (the C oracle engine behind the drop-in; CUDA == oracle is what tests/test_gpu_parity.py
proves) and only where the reference tree exists: the consumer is the reference's bytecode."""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import pytest

import helpers

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import ref_harness  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="needs the reference tree (consumer bytecode)")


def _consumer():
    ref = ref_harness.load_reference()   # installs the networkx-1.x shim BEFORE besst_b200.CreateGraph binds networkx
    import BESST.MakeScaffolds as MS
    import BESST.lp_solve as lps
    for name in ("Inf", "NaN"):   # numpy 2 dropped the aliases lp_solve.py:231 star-imports
        if not hasattr(lps, name):
            setattr(lps, name, getattr(np, name.lower()))
    return ref, MS


def _library(kind, seed, n_contigs=400, n_pairs=200000):
    from besst_b200 import synth
    mu, sigma, orient = {"pe": (550.0, 50.0, "fr"), "mp": (3000.0, 500.0, "rf")}[kind]
    lib = synth.make_library(n_contigs, n_pairs, orient, mu, sigma, 0.0, seed=seed)
    return lib.to_batch(), dict(orientation=orient, mean=mu, stddev=sigma, readlen=100)


def _scaffold_signature(o):
    def srows(d):
        return [(name, [(c.name, bool(c.direction), int(c.position), int(c.length)) for c in s.contigs], int(s.s_length))
                for name, s in d.items()]
    return {"Scaffolds": srows(o["Scaffolds"]), "small_scaffolds": srows(o["small_scaffolds"]),
            "Contigs": sorted(o["Contigs"]), "small_contigs": sorted(o["small_contigs"]),
            "G_prime": helpers.graph_signature(o["G_prime"]), "scaffold_indexer": o["param"].scaffold_indexer}


class _Pipeline(object):
    """runBESST's library loop (:143-231) around a `pe` callable: reference PE or the drop-in."""

    def __init__(self, ref, MS, use_dropin):
        self.ref, self.MS, self.use_dropin = ref, MS, use_dropin
        self.outdir = tempfile.mkdtemp(prefix="besst_consumer_")
        self.info = io.StringIO()
        self.Contigs, self.Scaffolds, self.small_contigs, self.small_scaffolds = {}, {}, {}, {}
        self.param = None

    def library(self, batch, opts, pass_number):
        ref = self.ref
        first = pass_number == 1
        if first:
            self.param = ref_harness.make_param(ref, opts, self.outdir, first_lib=True, pass_number=1)
        else:   # runBESST:144-158 re-assigns the per-library fields on the same object
            p = ref_harness.make_param(ref, opts, self.outdir, first_lib=False, pass_number=pass_number)
            for k in ("pass_number", "orientation", "mean_ins_size", "ins_size_threshold", "edgesupport", "read_len",
                      "std_dev_ins_size", "contig_threshold", "first_lib"):
                setattr(self.param, k, getattr(p, k))
        param = self.param
        param.information_file = self.info
        C_dict = {name: ref_harness.FakeSeq(n) for name, n in zip(batch.references, batch.lengths)} if first else {}
        with contextlib.redirect_stdout(io.StringIO()):
            if self.use_dropin:
                from oracle_engine import OracleEngine
                from besst_b200 import CreateGraph as CG, libmetrics
                from besst_b200.records import BatchFile
                eng = OracleEngine()
                bam = BatchFile(batch)
                param.contig_index = dict(zip(range(len(bam.references)), bam.references))
                libmetrics.get_metrics(bam, param, self.info, engine=eng)
                G, G_prime = CG.PE(self.Contigs, self.Scaffolds, self.info, C_dict, param, self.small_contigs,
                                   self.small_scaffolds, bam, engine=eng)
            else:
                bam = ref.pysam.Samfile(batch)
                param.contig_index = dict(zip(range(len(bam.references)), bam.references))
                ref.libmetrics.get_metrics(bam, param, self.info)
                G, G_prime = ref.CG.PE(self.Contigs, self.Scaffolds, self.info, C_dict, param, self.small_contigs,
                                       self.small_scaffolds, bam)
            after_pe = dict(G=helpers.graph_signature(G), G_prime=helpers.graph_signature(G_prime))
            self.MS.Algorithm(G, G_prime, self.Contigs, self.small_contigs, self.Scaffolds, self.small_scaffolds, self.info, param)
        return after_pe, _scaffold_signature(dict(Scaffolds=self.Scaffolds, small_scaffolds=self.small_scaffolds,
                                                  Contigs=self.Contigs, small_contigs=self.small_contigs,
                                                  G_prime=G_prime, param=param))


def _assert_same_outcome(a, b, label):
    pe_a, sc_a = a
    pe_b, sc_b = b
    helpers.assert_signature_equal(pe_b["G"], pe_a["G"], label=label + "/G after PE")
    helpers.assert_signature_equal(pe_b["G_prime"], pe_a["G_prime"], label=label + "/G_prime after PE")
    for key in ("Scaffolds", "small_scaffolds", "Contigs", "small_contigs", "scaffold_indexer"):
        assert sc_b[key] == sc_a[key], "%s: %s differs after MakeScaffolds.Algorithm" % (label, key)
    helpers.assert_signature_equal(sc_b["G_prime"], sc_a["G_prime"], label=label + "/G_prime after Algorithm")


@pytest.mark.parametrize("kind,seed", [("mp", 31), ("pe", 32)])
def test_makescaffolds_algorithm_on_dropin_graphs_one_library(kind, seed):
    ref, MS = _consumer()
    batch, opts = _library(kind, seed)
    want = _Pipeline(ref, MS, use_dropin=False).library(batch, opts, 1)
    got = _Pipeline(ref, MS, use_dropin=True).library(batch, opts, 1)
    _assert_same_outcome(want, got, kind)
    assert len(want[1]["Scaffolds"]) > 0


def test_two_library_sequence_pe_then_mp():
    """Library 2 starts from the scaffolds MakeScaffolds.Algorithm really built from library 1."""
    ref, MS = _consumer()
    lib1, opts1 = _library("pe", 41)
    lib2, opts2 = _library("mp", 41)   # same seed -> same contigs
    assert lib1.references == lib2.references
    a, b = _Pipeline(ref, MS, use_dropin=False), _Pipeline(ref, MS, use_dropin=True)
    _assert_same_outcome(a.library(lib1, opts1, 1), b.library(lib1, opts1, 1), "lib1")
    multi = [s for s in a.Scaffolds.values() if len(s.contigs) > 1]
    assert multi, "library 1 should have joined some contigs"
    assert any(not c.direction or c.position > 0 for s in multi for c in s.contigs)
    _assert_same_outcome(a.library(lib2, opts2, 2), b.library(lib2, opts2, 2), "lib2")


def test_dropin_graph_is_built_under_the_networkx_1x_api():
    """besst_b200.CreateGraph must not rely on networkx >= 2 attributes (graph.nodes[...] is a METHOD in 1.x)."""
    ref, MS = _consumer()
    batch, opts = _library("mp", 31, n_contigs=60, n_pairs=20000)
    p = _Pipeline(ref, MS, use_dropin=True)
    after_pe, _ = p.library(batch, opts, 1)
    import networkx
    assert networkx.Graph is sys.modules["nx1compat"].Graph
    assert callable(networkx.Graph().nodes) and isinstance(networkx.Graph().nodes(), list)
    assert all(n[2] is not None for n in after_pe["G_prime"]["nodes"])   # node attribute 'length' (:715-716)
