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


def update_info_gaps_batch(mean, sigma, read_len, sum_obs, nr_links, c1_len, c2_len, dValuesTable=None, engine=None):
    """The gap `MakeScaffolds.UpdateInfo` assigns between two joined scaffolds (MakeScaffolds.py:428-466), for
    arrays of edges at once: with a library sd and >= 5 links, two contigs longer than mean + 4 sd take the
    pre-calculated table (`PreCalcMLvaluesOfdLongContigs`, :68; a miss falls back to the bisection), two
    contigs longer than sd + read_len take `GapEstimator`, everything else the naive
    int((n * mean - sum_obs) / n).  All bisections of the batch run in ONE kernel launch.
    -> (avg_gap int64[n] before UpdateInfo's `<= 1 -> 1` clamp (:470-471), naive bool[n]: the entries the
    reference appends to param.gap_estimations)"""
    sum_obs = np.atleast_1d(np.asarray(sum_obs, dtype=np.float64))
    nr = np.broadcast_to(np.asarray(nr_links, dtype=np.float64), sum_obs.shape)
    c1 = np.broadcast_to(np.asarray(c1_len, dtype=np.float64), sum_obs.shape)
    c2 = np.broadcast_to(np.asarray(c2_len, dtype=np.float64), sum_obs.shape)
    data_observation = (nr * mean - sum_obs) / nr                      # :430
    mean_obs = sum_obs / nr                                            # :432
    fancy = bool(sigma) & (nr >= 5)                                     # :442
    long_pair = fancy & (c1 > mean + 4 * sigma) & (c2 > mean + 4 * sigma)          # :444
    mid_pair = fancy & ~long_pair & (c1 > sigma + read_len) & (c2 > sigma + read_len)   # :452
    naive = ~(long_pair | mid_pair)
    gap = np.trunc(data_observation).astype(np.int64)                  # int(data_observation) :455,462
    search = mid_pair.copy()
    if long_pair.any():
        keys = np.rint(data_observation[long_pair]).astype(np.int64)   # int(round(x, 0)): half to even, like Python 3
        table = dValuesTable or {}
        looked = np.array([table.get(int(k), None) for k in keys.tolist()], dtype=object)
        miss = np.array([v is None for v in looked], dtype=bool)
        idx = np.nonzero(long_pair)[0]
        gap[idx[~miss]] = np.array([int(v) for v in looked[~miss]], dtype=np.int64)
        search[idx[miss]] = True                                       # :448-449 KeyError -> bisection
    if search.any():
        g, _ = gap_estimator_batch(mean, sigma, read_len, mean_obs[search], c1[search], c2[search], engine=engine)
        gap[search] = g
    return gap, (~fancy) | (fancy & naive)


def lp_expected_means_batch(mean, sigma, read_len, mean_obs, c1_len, c2_len, engine=None):
    """`mean_obs + GapEstimator(...)` for every observation of a path (order_contigs.py:294-310, the constants of
    the gap LP) in one launch; call once per library parameter set (the PE-contamination links use
    contamination_mean / contamination_stddev, :299-300)."""
    mean_obs = np.atleast_1d(np.asarray(mean_obs, dtype=np.float64))
    g, _ = gap_estimator_batch(mean, sigma, read_len, mean_obs, c1_len, c2_len, engine=engine)
    return mean_obs + g


def lognormal_gap_estimator_batch(mu_ln, sigma_ln, read_len, samples, row_ptr, c1_len, c2_len=None, engine=None):
    """Drop-in for `mathstats.log_normal_param_est.GapEstimator` over many edges at once (CreateGraph.py:526,
    MakeScaffolds.py:425-426, order_contigs.py:304-306): samples[row_ptr[i]:row_ptr[i+1]] are the raw
    observations of edge i -- exactly the `observations` payload of the CSR the graph build returns.
    One warp per edge on the device.  -> gap int32[n]"""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    n = row_ptr.shape[0] - 1
    c1 = np.broadcast_to(np.asarray(c1_len, dtype=np.float64), (n,))
    c2 = c1 if c2_len is None else np.broadcast_to(np.asarray(c2_len, dtype=np.float64), (n,))
    return _engine(engine).gapest_lognormal_batch(mu_ln, sigma_ln, read_len, samples, row_ptr, c1, c2)


def LognormalGapEstimator(mu, sigma, read_length, samples, c1_len, c2_len=None, engine=None):
    """scalar form, same signature as mathstats.log_normal_param_est.GapEstimator"""
    samples = np.asarray(samples, dtype=np.int32)
    return int(lognormal_gap_estimator_batch(mu, sigma, read_length, samples, [0, samples.shape[0]], [c1_len],
                                             None if c2_len is None else [c2_len], engine=engine)[0])
