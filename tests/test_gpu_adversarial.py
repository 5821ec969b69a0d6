"""-m gpu: the CUDA graph build against the C oracle on the ADVERSARIAL random libraries of tests/test_differential_fuzz.py
(records drawn independently of each other, contigs around the threshold, unmapped reads / mates, mapq 0, duplicates,
secondary alignments, first and later libraries, every option combination) -- the inputs on which the CPU suite checks the
oracle against the reference's bytecode."""
import numpy as np
import pytest

import helpers
import oracle_lib
from besst_b200 import abi
from test_differential_fuzz import random_batch


@pytest.mark.gpu
def test_cuda_equals_oracle_on_adversarial_random_libraries(cuda_engine):
    n_graphs = 0
    for seed in range(64):
        rng = np.random.default_rng(seed)
        batch = random_batch(rng)
        mu, sd = float(rng.choice([400, 1500, 3000])), float(rng.choice([40, 150, 400]))
        params = abi.make_params("fr" if rng.random() < 0.5 else "rf", 0 if rng.random() < 0.2 else 11, 100.0, mu, sd, mu + 6 * sd,
                                 detect_duplicate=bool(rng.random() < 0.8), extend_paths=bool(rng.random() < 0.7),
                                 no_score=bool(rng.random() < 0.3))
        if rng.random() < 0.4:
            objs = helpers.later_library_objects(batch.references, batch.lengths, mu + 4 * sd, seed=int(rng.integers(1, 1000)))
        else:
            objs = helpers.first_library_objects(batch.references, batch.lengths, mu + 4 * sd)
        table = helpers.table_for(batch, objs)
        if table.n_scaffolds == 0:
            continue
        want, _, _, consistent = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
        assert consistent
        got = cuda_engine.graph_build(table, params, batch)
        helpers.assert_graph_equal(got, want, label="adversarial seed %d" % seed)
        n_graphs += want.n_edges > 0
    assert n_graphs > 40
