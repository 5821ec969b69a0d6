"""The drop-in entry points (besst_b200.libmetrics.get_metrics + besst_b200.CreateGraph.PE)
against golden dumps minted from the reference's own bytecode (oracle/make_golden.py,
tests/golden/*.json.gz): graphs with node/edge order and every attribute, `param`,
object dicts with coverages, counter lines.

  not gpu: the host mirror with the C oracle behind it (pins the oracle and the host logic)
  gpu:     the same with the CUDA engine behind it, through the C ABI
Integers bit-exact; scores to 1e-6 relative (north-star tolerance); library-metric
floats to 1e-9 (histogram-order summation, see besst_metrics.cu)."""
import gzip
import json
import os
import re

import pytest

import helpers
from besst_b200 import synth
from besst_b200.records import RecordBatch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(f[:-8] for f in os.listdir(GOLDEN) if f.endswith(".json.gz"))

COUNTER_PATTERNS = {
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

_inputs = {}


def load_input(name):
    if name not in _inputs:
        if name.startswith("sequence:"):   # the libraries of the PE -> MP sequence case (oracle/make_golden.py:sequence_library)
            mu, sigma, orient = {"pe": (550.0, 50.0, "fr"), "mp": (3000.0, 500.0, "rf")}[name.split(":")[1]]
            _inputs[name] = synth.make_library(400, 200000, orient, mu, sigma, 0.0, seed=41).to_batch()
        elif name == "testset1_full":
            _inputs[name] = RecordBatch.load(os.path.join(GOLDEN, "testset1_full.npz"))
        elif name in ("testset1", "testset2"):
            _inputs[name] = RecordBatch.load(os.path.join(GOLDEN, name + "_head.npz"))
        else:
            _inputs[name] = synth.make_config(name).to_batch()
    return _inputs[name]


def load_golden(case):
    with gzip.open(os.path.join(GOLDEN, case + ".json.gz"), "rb") as fh:
        return json.loads(fh.read().decode())


def contig_threshold_for(opts):
    if opts.get("minsize"):
        return opts["minsize"]
    if opts.get("extendpaths", True):
        return opts["mean"] + 4 * opts["stddev"]
    return opts["mean"] + (opts["stddev"] / float(opts["mean"])) * opts["stddev"]


def check_case(case, engine):
    g = load_golden(case)
    batch = load_input(g["input"])
    assert len(batch) == g["n_records"]
    opts = g["options"]
    state = None
    if g.get("state") is not None:     # the state a REAL scaffolding pass of the reference left behind
        state = helpers.state_from_snapshot(g["state"])
    elif g["later_seed"] is not None:
        state = helpers.state_for_later_library(batch, contig_threshold_for(opts), g["later_seed"])
    out = helpers.run_dropin(batch, opts, engine, state=state)
    helpers.assert_param_equal(out["param"], g["param"], label=case)
    helpers.assert_signature_equal(out["G"], g["G"], label=case + "/G")
    helpers.assert_signature_equal(out["G_prime"], g["G_prime"], label=case + "/G_prime")
    for key in ("Contigs", "small_contigs", "Scaffolds", "small_scaffolds"):
        assert out["objects"][key] == g["objects"][key], "%s: %s differs" % (case, key)
    for key, pat in COUNTER_PATTERNS.items():
        want = g["counters"].get(key)
        m = re.search(pat, out["information"])
        got = int(m.group(1)) if m else None
        assert got == want, "%s: counter line %s: %r != %r" % (case, key, got, want)


@pytest.mark.parametrize("case", CASES)
def test_dropin_with_oracle_matches_reference(case):
    from oracle_engine import OracleEngine
    check_case(case, OracleEngine())


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_dropin_with_cuda_matches_reference(case, cuda_engine):
    check_case(case, cuda_engine)
