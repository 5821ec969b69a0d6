"""ORACLE / TEST INFRASTRUCTURE -- mint the golden fixtures under tests/golden/.

Runs the reference's own, unmodified bytecode (oracle/ref_harness.py:
/root/reference/BESST/{libmetrics,CreateGraph}.py with pysam / networkx-1.x /
mathstats stand-ins) on

  * committed slices of the reference's testdata/testset1/mapped.bam and testset2/mapped.bam (the
    first N records, decoded by besst_b200/bamio.py) with the two Travis command lines
    (.travis.yml:14-15) and without -m/-s so libmetrics is exercised;
  * seeded synthetic libraries (besst_b200/synth.py) as first and later libraries,
    with the option combinations the record loop branches on,

and stores, per case, the canonical dump of what `get_metrics` + `PE` leave
behind: (G, G_prime) with node/edge order and every attribute, `param`, the four
object dicts with coverages, and the counter lines of Statistics.txt.

Only runs where /root/reference exists (this container).  Usage:
    python oracle/make_golden.py            # writes tests/golden/*
"""
from __future__ import annotations

import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import helpers  # noqa: E402
import ref_harness  # noqa: E402
from besst_b200 import bamio, synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
TESTSET1_RECORDS = 300000
TESTSET2_RECORDS = 250000

# name -> (input, options, later-library seed or None, run libmetrics)
CASES = {
    "testset1_travis": ("testset1", dict(orientation="fr", mean=4000, stddev=500, minsize=3000, threshold=6000), None, True),
    "testset1_travis_no_score": ("testset1", dict(orientation="fr", mean=4000, stddev=500, minsize=3000, threshold=6000, no_score=True), None, True),
    "testset1_auto": ("testset1", dict(orientation="fr"), None, True),
    # BASELINE config 1 in full: all 1,999,958 records of testdata/testset1/mapped.bam with the Travis command lines
    "testset1_full_travis": ("testset1_full", dict(orientation="fr", mean=4000, stddev=500, minsize=3000, threshold=6000), None, True),
    "testset1_full_travis_no_score": ("testset1_full", dict(orientation="fr", mean=4000, stddev=500, minsize=3000, threshold=6000, no_score=True), None, True),
    "testset2_auto": ("testset2", dict(orientation="fr"), None, True),
    "testset2_given": ("testset2", dict(orientation="fr", mean=2800, stddev=350, minsize=4600, threshold=4500), None, True),
    "small_pe_auto": ("small_pe", dict(orientation="fr"), None, True),
    "small_mp_auto": ("small_mp", dict(orientation="rf"), None, True),
    "small_mp_cont_auto": ("small_mp_cont", dict(orientation="rf"), None, True),
    "small_mp_given": ("small_mp", dict(orientation="rf", mean=3000.0, stddev=500.0, readlen=100), None, True),
    "small_pe_later": ("small_pe", dict(orientation="fr", mean=550.0, stddev=50.0, readlen=100), 7, True),
    "small_mp_later_nodup": ("small_mp", dict(orientation="rf", mean=3000.0, stddev=500.0, readlen=100, duplicate=False), 11, True),
    "small_mp_no_extend": ("small_mp", dict(orientation="rf", mean=3000.0, stddev=500.0, readlen=100, extendpaths=False), None, True),
    "small_pe_no_score_later": ("small_pe", dict(orientation="fr", mean=550.0, stddev=50.0, readlen=99.37, no_score=True), 5, True),
    "tiny_mapq0": ("tiny", dict(orientation="fr", mean=550.0, stddev=50.0, readlen=100, min_mapq=0, edgesupport=3), None, True),
}


def load_input(name):
    if name == "testset1_full":
        from besst_b200.records import RecordBatch
        return RecordBatch.load(os.path.join(GOLDEN, "testset1_full.npz"))
    if name in ("testset1", "testset2"):
        path = os.path.join(GOLDEN, name + "_head.npz")
        from besst_b200.records import RecordBatch
        return RecordBatch.load(path)
    return synth.make_config(name).to_batch()


def make_testset1_fixture():
    bam = os.path.join(ref_harness.REFERENCE_ROOT, "testdata", "testset1", "mapped.bam")
    batch = bamio.read_bam(bam, max_records=TESTSET1_RECORDS)
    batch.save(os.path.join(GOLDEN, "testset1_head.npz"))
    bamio.read_bam_native(bam).save_compact(os.path.join(GOLDEN, "testset1_full.npz"))   # every record (native reader == read_bam, tests/test_bamio.py)
    bam2 = os.path.join(ref_harness.REFERENCE_ROOT, "testdata", "testset2", "mapped.bam")
    bamio.read_bam(bam2, max_records=TESTSET2_RECORDS).save(os.path.join(GOLDEN, "testset2_head.npz"))
    return batch


def contig_threshold_for(opts):
    if opts.get("minsize"):
        return opts["minsize"]
    return opts["mean"] + 4 * opts["stddev"] if opts.get("extendpaths", True) else opts["mean"] + (opts["stddev"] / float(opts["mean"])) * opts["stddev"]


def run_case(name):
    inp, opts, later_seed, run_lm = CASES[name]
    batch = load_input(inp)
    state = None
    if later_seed is not None:
        state = helpers.state_for_later_library(batch, contig_threshold_for(opts), later_seed)
    out = ref_harness.run_reference(batch, opts, state=state, run_libmetrics=run_lm)
    objs = out["objects"]
    golden = {
        "case": name, "input": inp, "options": opts, "later_seed": later_seed,
        "G": helpers.graph_signature(objs["G"]), "G_prime": helpers.graph_signature(objs["G_prime"]),
        "param": helpers.param_signature(objs["param"]),
        "objects": helpers.object_signature(objs["Contigs"], objs["Scaffolds"], objs["small_contigs"], objs["small_scaffolds"]),
        "counters": out["counters"],
        "n_records": len(batch),
    }
    return golden


SEQUENCE_SEED = 41


def sequence_library(kind):
    """the two libraries of the PE -> MP sequence case: same contigs (same seed), different inserts"""
    mu, sigma, orient = {"pe": (550.0, 50.0, "fr"), "mp": (3000.0, 500.0, "rf")}[kind]
    return synth.make_library(400, 200000, orient, mu, sigma, 0.0, seed=SEQUENCE_SEED).to_batch(), dict(orientation=orient, mean=mu, stddev=sigma, readlen=100)


def run_sequence_case():
    """Library 2 of a run whose library 1 went through the reference's REAL scaffolding pass: get_metrics + PE +
    MakeScaffolds.Algorithm (runBESST:168-199) on a PE library, then the golden of get_metrics + PE of an MP library that
    starts from the scaffolds that pass built (CleanObjects :788-810, multi-contig PosDir cases :1031-1048).  The state
    between the two libraries is stored in the golden, so the test needs no reference."""
    import io
    import numpy as np
    ref = ref_harness.load_reference()
    import BESST.MakeScaffolds as MS
    import BESST.lp_solve as lps
    for name in ("Inf", "NaN"):   # numpy 2 dropped the aliases lp_solve.py star-imports
        if not hasattr(lps, name):
            setattr(lps, name, getattr(np, name.lower()))
    lib1, opts1 = sequence_library("pe")
    lib2, opts2 = sequence_library("mp")
    out1 = ref_harness.run_reference(lib1, opts1)
    o = out1["objects"]
    MS.Algorithm(o["G"], o["G_prime"], o["Contigs"], o["small_contigs"], o["Scaffolds"], o["small_scaffolds"], io.StringIO(), o["param"])
    snap = helpers.state_snapshot(o["Contigs"], o["Scaffolds"], o["small_contigs"], o["small_scaffolds"], o["param"])
    assert any(len(members) > 1 for _, members, _ in snap["Scaffolds"])
    out2 = ref_harness.run_reference(lib2, opts2, state=helpers.state_from_snapshot(snap))
    objs = out2["objects"]
    return {
        "case": "sequence_mp_after_real_pass", "input": "sequence:mp", "options": opts2, "later_seed": None, "state": snap,
        "G": helpers.graph_signature(objs["G"]), "G_prime": helpers.graph_signature(objs["G_prime"]),
        "param": helpers.param_signature(objs["param"]),
        "objects": helpers.object_signature(objs["Contigs"], objs["Scaffolds"], objs["small_contigs"], objs["small_scaffolds"]),
        "counters": out2["counters"], "n_records": len(lib2),
    }


def main():
    if not ref_harness.reference_available():
        sys.exit("reference tree not found: golden fixtures can only be minted where /root/reference exists")
    os.makedirs(GOLDEN, exist_ok=True)
    make_testset1_fixture()
    for name in CASES:
        g = run_case(name)
        path = os.path.join(GOLDEN, name + ".json.gz")
        with gzip.GzipFile(path, "wb", mtime=0) as fh:
            fh.write(json.dumps(g, sort_keys=True).encode())
        print("%-28s G %5d edges  G_prime %6d edges  %s" % (name, len(g["G"]["edges"]), len(g["G_prime"]["edges"]),
                                                            {k: g["counters"][k] for k in ("count", "duplicates", "fishy")}))
    g = run_sequence_case()
    with gzip.GzipFile(os.path.join(GOLDEN, g["case"] + ".json.gz"), "wb", mtime=0) as fh:
        fh.write(json.dumps(g, sort_keys=True).encode())
    print("%-28s G %5d edges  G_prime %6d edges  %d multi-contig scaffolds in the state" % (
        g["case"], len(g["G"]["edges"]), len(g["G_prime"]["edges"]), sum(len(m) > 1 for _, m, _ in g["state"]["Scaffolds"])))


if __name__ == "__main__":
    main()
