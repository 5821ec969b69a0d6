"""ORACLE / TEST INFRASTRUCTURE -- an engine object with the same interface as
besst_b200.engine.CudaEngine, backed by the sequential C oracle.  Lets the CPU
test-suite exercise the host-side mirror (besst_b200/CreateGraph.py,
libmetrics.py) end to end without a GPU.  Tests only."""
from __future__ import annotations

import oracle_lib


class OracleEngine(object):
    name = "oracle"

    def graph_build(self, table, params, batch, view=False):
        res, tuples, fishy, consistent = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
        assert consistent, "oracle: G and G_prime disagree on large-large edges"
        self.last_tuples = tuples
        return res

    def libmetrics(self, table_rows, params, batch, ref_lengths, want_isize, cap=1 << 20):
        return oracle_lib.libmetrics(table_rows, params, batch, ref_lengths, want_isize, cap)

    def gapest_batch(self, params, mean_obs, len1, len2):
        return oracle_lib.gapest_batch(params, mean_obs, len1, len2)

    def gapest_lognormal_batch(self, mu, sigma, read_len, samples, row_ptr, len1, len2):
        return oracle_lib.gapest_lognormal_batch(mu, sigma, read_len, samples, row_ptr, len1, len2)
