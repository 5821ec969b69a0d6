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
