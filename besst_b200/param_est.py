"""Drop-in for `mathstats.normaldist.truncatedskewed.param_est` as BESST calls
it: `GapEstimator` (CreateGraph.py:537; MakeScaffolds.py:449,453;
order_contigs.py:300,308; pathgaps.py:108,204) and `tr_sk_std_dev`
(CreateGraph.py:555), plus the batched forms the scalar ones are built on.

The arithmetic runs in the k_gapest_batch CUDA kernel (one quad of lanes per
item, besst_gapest_batch in include/besst_b200.h).  A scalar call is a batch of
one: callers with many contig pairs (UpdateInfo, the per-path LP set-up) should
use the *_batch functions.  There is no CPU fallback.
"""
from __future__ import annotations

import sys

import numpy as np

from . import abi


def _engine(engine):
    if engine is not None:
        return engine
    from .engine import default_engine
    return default_engine()


def _params(mean, sigma, read_len, erf_variant):
    return abi.make_params("fr", 0, read_len, mean, sigma, 0.0, erf_variant=erf_variant)


def gap_estimator_batch(mean, sigma, read_len, mean_obs, c1_len, c2_len=None, engine=None,
                        erf_variant=abi.ERF_AS7126):
    """ML gap for every (mean_obs[i], c1_len[i], c2_len[i]) under one library
    (mean, sigma, read_len).  -> (gap int32[n], tr_sk_std_dev at that gap float64[n])"""
    mean_obs = np.atleast_1d(np.asarray(mean_obs, dtype=np.float64))
    c1 = np.broadcast_to(np.asarray(c1_len), mean_obs.shape)
    c2 = c1 if c2_len is None else np.broadcast_to(np.asarray(c2_len), mean_obs.shape)
    if (np.asarray(c1) <= 0).any() or (np.asarray(c2) <= 0).any():
        sys.stderr.write('ERROR! Gap estimation on contigs with negative length\n')
    return _engine(engine).gapest_batch(_params(mean, sigma, read_len, erf_variant), mean_obs, c1, c2)


def GapEstimator(mean, sigma, read_length, mean_obs, c1_len, c2_len=None, engine=None):
    gap, _ = gap_estimator_batch(mean, sigma, read_length, [mean_obs], [c1_len],
                                 None if c2_len is None else [c2_len], engine=engine)
    return int(gap[0])


def tr_sk_std_dev_batch(mean, sigma, read_len, c1_len, c2_len, gap, engine=None, erf_variant=abi.ERF_AS7126):
    gap = np.atleast_1d(np.asarray(gap, dtype=np.float64))
    c1 = np.broadcast_to(np.asarray(c1_len), gap.shape)
    c2 = np.broadcast_to(np.asarray(c2_len), gap.shape)
    return _engine(engine).trsk_sd_batch(_params(mean, sigma, read_len, erf_variant), gap, c1, c2)


def tr_sk_std_dev(mean, stdDev, readLen, c1Len, c2Len, d, engine=None):
    return float(tr_sk_std_dev_batch(mean, stdDev, readLen, [c1Len], [c2Len], [d], engine=engine)[0])


def func_of_d_batch(mean, sigma, read_len, d, c1_len, c2_len, engine=None, erf_variant=abi.ERF_AS7126):
    """d + sigma^2 g'(d)/g(d) (mathstats funcDGeneral) for arrays of d and contig lengths."""
    d = np.atleast_1d(np.asarray(d, dtype=np.float64))
    c1 = np.broadcast_to(np.asarray(c1_len), d.shape)
    c2 = np.broadcast_to(np.asarray(c2_len), d.shape)
    return _engine(engine).func_of_d_batch(_params(mean, sigma, read_len, erf_variant), d, c1, c2)


def PreCalcMLvaluesOfdLongContigs(mean, stdDev, readLen, engine=None):
    """Drop-in for mathstats' table {rounded left-hand side of the ML equation: gap d} for two long contigs
    (c1 = c2 = mean + 4 stdDev), built once per library by MakeScaffolds.py:68 and looked up in UpdateInfo.
    All d in [int(-4 sd), int(mean + 2 sd - 2 r)] are evaluated in one kernel launch; the dictionary is
    filled in the reference's order (holes between consecutive rounded values take the larger d)."""
    d_upper = int(mean + 2 * stdDev - 2 * readLen)
    d_lower = int(-4 * stdDev)
    table = {}
    if d_upper < d_lower:
        return table
    ds = np.arange(d_lower, d_upper + 1, dtype=np.float64)
    # c1 = c2 = mean + 4 stdDev as a float: the call site passes ESTIMATED library parameters
    # (mu_adj, sigma_adj of get_metrics), so the length is never integral; lengths are fp64 across the ABI
    c = float(mean + 4 * stdDev)
    f = func_of_d_batch(mean, stdDev, readLen, ds, c, c, engine=engine)
    prev_obs = d_lower
    for d, func_of_d in zip(range(d_lower, d_upper + 1), f.tolist()):
        obs = int(round(func_of_d, 0))
        table[obs] = d
        if abs(obs - prev_obs) > 1:
            for i in range(abs(obs - prev_obs)):
                table[prev_obs + i + 1] = d
        prev_obs = obs
    return table
