"""Pins for the oracle's restated arithmetic (mathstats is not vendored by the
reference: PARITY UNPINNED at that boundary).  What can be pinned:
the closed forms against numerical quadrature of the published GapEst model, the
C restatement against the Python restatement bit for bit, KS against scipy, the
inverse normal CDF against scipy within its published error, and e_nr_links
against the reference module's value recorded in SURVEY.md 8c."""
import ctypes
import math
import os
import sys

import numpy as np
import pytest

import oracle_lib
from besst_b200 import abi, e_nr_links, normal

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle", "mathstats_restated"))
from mathstats.normaldist import normal as ms_normal  # noqa: E402
from mathstats.normaldist.truncatedskewed import param_est as ms_pe  # noqa: E402


def model_moments(d, mean, sd, c1, c2, r, n=400001):
    """g(d), E[x], E[x^2] of the density w(x-d) phi(x) by quadrature."""
    c_min, c_max = min(c1, c2), max(c1, c2)
    lo, hi = d + 2 * r - 1, d + c_min + c_max + 1
    x = np.linspace(lo, hi, n)
    o = x - d
    w = np.minimum(np.minimum(o - 2 * r + 1, c_min - r + 1), c_min + c_max - o + 1)
    w = np.maximum(w, 0)
    phi = np.exp(-(x - mean) ** 2 / (2 * sd * sd)) / (math.sqrt(2 * math.pi) * sd)
    f = w * phi
    g = np.trapezoid(f, x)
    return g, np.trapezoid(f * x, x) / g, np.trapezoid(f * x * x, x) / g


@pytest.mark.parametrize("mean,sd,r,c1,c2,d", [
    (3000.0, 500.0, 100.0, 6000, 9000, 400.0), (3000.0, 500.0, 100.0, 2500, 12000, -150.0),
    (550.0, 50.0, 100.0, 1000, 700, 120.0), (8000.0, 1200.0, 100.0, 20000, 5000, 2500.0)])
def test_closed_forms_match_quadrature(mean, sd, r, c1, c2, d, monkeypatch):
    monkeypatch.setattr(ms_pe, "erf", math.erf)   # the closed form itself; A&S erf adds ~1e-7
    c_min, c_max = min(c1, c2), max(c1, c2)
    g, gp, gb = ms_pe._g_and_gprime(d, mean, sd, c_min, c_max, r)
    gq, ex, ex2 = model_moments(d, mean, sd, c1, c2, r)
    assert g == pytest.approx(gq, rel=1e-6)
    # E[x] = mu - sigma^2 g'/g  <=>  the ML equation mu - mean_obs = d + sigma^2 g'/g
    assert mean - sd * sd * gp / g == pytest.approx(ex, rel=1e-6)
    sd_model = ms_pe.tr_sk_std_dev(mean, sd, r, c1, c2, d)
    assert sd_model == pytest.approx(math.sqrt(ex2 - ex * ex), rel=1e-5)


def test_gap_estimator_inverts_the_model_mean():
    mean, sd, r, c1, c2 = 3000.0, 500.0, 100.0, 7000, 4000
    for true_gap in (-200, 0, 350, 1200, 2300):
        _, ex, _ = model_moments(true_gap, mean, sd, c1, c2, r)
        est = ms_pe.GapEstimator(mean, sd, r, ex - true_gap, c1, c2)
        assert abs(est - true_gap) <= 1


def test_c_oracle_equals_python_restatement_bitwise():
    rng = np.random.default_rng(5)
    L = oracle_lib.lib()
    for _ in range(300):
        mean = float(rng.choice([550.0, 3000.0, 3200.49, 8000.0]))
        sd = mean * float(rng.uniform(0.05, 0.25))
        r = float(rng.choice([100.0, 99.37, 150.0]))
        c1, c2 = int(rng.integers(600, 40000)), int(rng.integers(600, 40000))
        mo = float(rng.uniform(2 * r, mean + 2 * sd))
        g_py = ms_pe.GapEstimator(mean, sd, r, mo, c1, c2)
        g_c = L.besst_oracle_gap_estimator(mean, sd, r, mo, float(c1), float(c2), abi.ERF_AS7126)
        assert g_py == g_c
        s_py = ms_pe.tr_sk_std_dev(mean, sd, r, c1, c2, g_py)
        s_c = L.besst_oracle_tr_sk_std_dev(mean, sd, r, float(c1), float(c2), float(g_py), abi.ERF_AS7126)
        assert s_py == s_c or (math.isnan(s_py) and math.isnan(s_c))


def test_max_obs_distr():
    from scipy.stats import norm
    L = oracle_lib.lib()
    for n in (2, 10, 1000, 50000, 1000000):
        k = ms_normal.MaxObsDistr(n, 0.95)
        assert k == L.besst_oracle_max_obs_distr(float(n), 0.95) == normal.MaxObsDistr(n, 0.95)
        assert abs(k - norm.ppf(0.95 ** (1.0 / n))) < 4.5e-4   # A&S 26.2.23 error bound


def test_erf_variant_error_bound():
    for x in np.linspace(-4, 4, 161):
        assert abs(ms_normal.erf(float(x)) - math.erf(float(x))) <= 1.5e-7


def test_ks_matches_scipy():
    from scipy.stats import ks_2samp
    rng = np.random.default_rng(3)
    L = oracle_lib.lib()
    assert L.besst_oracle_ks_2samp(np.array([1, 2, 3, 4, 5.0]).ctypes.data, 5,
                                   np.array([1.5, 2.5, 3.5, 4.5, 9]).ctypes.data, 5) == 0.20000000000000007
    for n in (1, 2, 5, 37, 400):
        a = np.ascontiguousarray(rng.integers(0, 50, n).astype(np.float64) - 24.3)
        b = np.ascontiguousarray(rng.integers(0, 50, n).astype(np.float64) - 25.1)
        got = L.besst_oracle_ks_2samp(a.ctypes.data, n, b.ctypes.data, n)
        # the formula of the scipy the reference pins (requirements.txt:4, scipy==1.0.0):
        # max |searchsorted(d1, all, 'right')/n1 - searchsorted(d2, all, 'right')/n2|
        d1, d2 = np.sort(a), np.sort(b)
        pooled = np.concatenate([d1, d2])
        want = np.max(np.abs(np.searchsorted(d1, pooled, side="right") / (1.0 * n)
                             - np.searchsorted(d2, pooled, side="right") / (1.0 * n)))
        assert got == want
        # the installed scipy (1.18) builds the ECDF values with linspace: equal to rounding
        assert got == pytest.approx(ks_2samp(a, b).statistic, rel=0, abs=1e-12)


def test_expected_links_known_answer():
    # value of the reference's own BESST/e_nr_links.py in this image
    v = e_nr_links.ExpectedLinks(1e5, 1e5, 3300, e_nr_links.Param(3000, 500, 50, 100, 0))
    assert v == 10.41443382345824
    ref_root = os.environ.get("BESST_REFERENCE_ROOT", "/root/reference")
    if os.path.isfile(os.path.join(ref_root, "BESST", "e_nr_links.py")):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ref_e_nr_links", os.path.join(ref_root, "BESST", "e_nr_links.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        for args in ((1e5, 1e5, 3300, (3000, 500, 50, 100, 0)), (5000, 80000, 120, (550, 50, 21.3, 99.37, 0)),
                     (700, 900, -40, (3200.5, 499.2, 35.0, 100, 0))):
            a = e_nr_links.ExpectedLinks(args[0], args[1], args[2], e_nr_links.Param(*args[3]))
            b = mod.ExpectedLinks(args[0], args[1], args[2], mod.Param(*args[3]))
            assert a == b


def test_update_info_gap_batch_follows_the_reference_guards():
    """MakeScaffolds.UpdateInfo's gap choice (:428-466) replayed edge by edge with the restated mathstats against the
    batched form (one launch for all bisections); estimated, non-integral library parameters as get_metrics leaves them."""
    GC = ms_pe
    from besst_b200 import param_est
    from oracle_engine import OracleEngine
    mean, sd, r = 2987.4142135, 512.7182818, 99.37
    table = GC.PreCalcMLvaluesOfdLongContigs(mean, sd, r)
    rng = np.random.default_rng(8)
    n = 400
    nr = rng.integers(1, 60, n)
    c1 = rng.choice([300, 700, 2000, 6000, 9000], n).astype(np.float64)
    c2 = rng.choice([450, 800, 5100, 5200, 12000], n).astype(np.float64)
    mean_obs = rng.uniform(300.0, 4200.0, n)       # includes observations outside the table's range
    sum_obs = np.floor(mean_obs * nr)
    want, naive = [], []
    for i in range(n):
        data_observation = (nr[i] * mean - sum_obs[i]) / float(nr[i])
        mo = sum_obs[i] / float(nr[i])
        is_naive = False
        if sd and nr[i] >= 5:
            if c1[i] > mean + 4 * sd and c2[i] > mean + 4 * sd:
                try:
                    g = table[int(round(data_observation, 0))]
                except KeyError:
                    g = GC.GapEstimator(mean, sd, r, mo, c1[i], c2[i])
            elif c1[i] > sd + r and c2[i] > sd + r:
                g = GC.GapEstimator(mean, sd, r, mo, c1[i], c2[i])
            else:
                g, is_naive = int(data_observation), True
        else:
            g, is_naive = int(data_observation), True
        want.append(g)
        naive.append(is_naive)
    got, got_naive = param_est.update_info_gaps_batch(mean, sd, r, sum_obs, nr, c1, c2, table, engine=OracleEngine())
    assert got.tolist() == want and got_naive.tolist() == naive
    assert 20 < sum(naive) < n - 50
    lp = param_est.lp_expected_means_batch(mean, sd, r, mean_obs[:50], c1[:50] + 1000, c2[:50] + 1000, engine=OracleEngine())
    assert lp.tolist() == [mean_obs[i] + GC.GapEstimator(mean, sd, r, mean_obs[i], c1[i] + 1000, c2[i] + 1000) for i in range(50)]


# ---- lognormal GapEstimator (mathstats.log_normal_param_est, restated from the model) ----------------------------------
from mathstats import log_normal_param_est as ms_ln  # noqa: E402


def _trapezoid_weight(o, c_min, c_max, r):
    return np.clip(np.minimum(np.minimum(o - 2 * r + 1, c_min - r + 1), c_min + c_max + 1 - o), 0, None)


@pytest.mark.parametrize("d,c1,c2", [(500, 6000, 8000), (1500, 2500, 9000), (-40, 3000, 3000), (2500, 20000, 30000)])
def test_lognormal_g_closed_form_equals_quadrature(d, c1, c2):
    mu, sigma, r = 8.0, 0.25, 100
    c_min, c_max = min(c1, c2), max(c1, c2)
    xs = np.linspace(1, 60000, 2000001)
    f = np.exp(-(np.log(xs) - mu) ** 2 / (2 * sigma ** 2)) / (xs * sigma * np.sqrt(2 * np.pi))
    quad = np.trapezoid(_trapezoid_weight(xs - d, c_min, c_max, r) * f, xs)
    assert ms_ln.g_of_d(d, mu, sigma, c_min, c_max, r) == pytest.approx(quad, rel=1e-9)
    L = oracle_lib.lib()
    L.besst_oracle_lognormal_g.restype = ctypes.c_double
    L.besst_oracle_lognormal_g.argtypes = [ctypes.c_double] * 6
    assert L.besst_oracle_lognormal_g(d, mu, sigma, c_min, c_max, r) == ms_ln.g_of_d(d, mu, sigma, c_min, c_max, r)


def _simulate(rng, mu, sigma, r, c1, c2, d, n):
    c_min, c_max = min(c1, c2), max(c1, c2)
    out = []
    while len(out) < n:
        o = np.rint(rng.lognormal(mu, sigma, 4000) - d)
        acc = rng.random(4000) * (c_min - r + 1) < _trapezoid_weight(o, c_min, c_max, r)
        out += [int(v) for v in o[acc]]
    return out[:n]


def test_lognormal_estimator_recovers_the_simulated_gap_and_c_equals_python():
    rng = np.random.default_rng(1)
    mu, sigma, r = 8.0, 0.25, 100
    samples, row_ptr, l1, l2, truth = [], [0], [], [], []
    for (c1, c2, d, n) in ((6000, 8000, 500, 300), (2500, 9000, 1500, 200), (3000, 3000, -40, 400), (20000, 30000, 2500, 60)):
        for _ in range(12):
            samples += _simulate(rng, mu, sigma, r, c1, c2, d, n)
            row_ptr.append(len(samples)); l1.append(c1); l2.append(c2); truth.append(d)
    got = oracle_lib.gapest_lognormal_batch(mu, sigma, r, samples, row_ptr, l1, l2)
    want = [ms_ln.GapEstimator(mu, sigma, r, samples[row_ptr[i]:row_ptr[i + 1]], l1[i], l2[i]) for i in range(len(l1))]
    assert got.tolist() == want                       # C restatement == Python restatement, bit for bit
    err = got.reshape(4, 12).mean(axis=1) - np.array([500, 1500, -40, 2500])
    naive = np.array([np.exp(mu + sigma ** 2 / 2) - np.mean(samples[row_ptr[12 * k]:row_ptr[12 * k + 12]]) for k in range(4)]) - np.array([500, 1500, -40, 2500])
    # within 4 standard errors of the truth (sd of one estimate ~ 41 / 108 / 40 / 226 bp), where the naive estimate is 300-900 bp off
    assert np.all(np.abs(err) < np.array([48, 125, 48, 260])) and np.all(np.abs(err[[0, 1, 3]]) < 0.3 * np.abs(naive[[0, 1, 3]]))


def test_lognormal_scoring_branch_vectorised_equals_scalar_replay():
    """besst_b200.CreateGraph.lognormal_rescore against an edge-by-edge replay of the reference's branch
    (CreateGraph.py:485-493,523-531,542-614) with the restated estimator."""
    import helpers
    from oracle_engine import OracleEngine
    from besst_b200 import CreateGraph as CG, synth
    lib = synth.make_config("small_mp")
    batch = lib.to_batch()
    params = abi.make_params("rf", 11, 100.0, 3000.0, 500.0, 6000.0)
    table = helpers.table_for(batch, helpers.first_library_objects(batch.references, batch.lengths, 5000.0))
    eng = OracleEngine()
    res = eng.graph_build(table, params, batch)
    base_gap, base_score = res.gap.copy(), res.score.copy()

    class P(object):
        mean_ins_size, read_len = 3000.0, 100.0
        lognormal_sigma = 0.17
        lognormal_mean = math.log(3000.0) - 0.17 ** 2 / 2
        empirical_distribution = {x: math.exp(-((x - 3000.0) / 500.0) ** 2 / 2) for x in range(200, 6001)}
    CG.lognormal_rescore(res, table, P, eng)
    emp, max_isize = P.empirical_distribution, 6000
    cond = CG.get_conditional_stddevs(list(range(0, int(max_isize * 0.8), max_isize // 50)), emp, max_isize)
    changed = 0
    for e in np.nonzero(res.flags & abi.EDGE_SCORED)[0].tolist():
        n = int(res.nr_links[e])
        len1, len2 = float(table.scaffold_lengths[res.edge_u[e] >> 1]), float(table.scaffold_lengths[res.edge_v[e] >> 1])
        samples = (res.obs_u[res.row_ptr[e]:res.row_ptr[e + 1]].astype(np.int64) + res.obs_v[res.row_ptr[e]:res.row_ptr[e + 1]]).tolist()
        if 2 * 500.0 < len1 and 2 * 500.0 < len2:
            gap = ms_ln.GapEstimator(P.lognormal_mean, P.lognormal_sigma, P.read_len, samples, len1, c2_len=len2)
            gap = min(gap, len(cond) - 1)
            sd0 = cond[int(gap)] if gap > 0 else cond[0]
        else:
            gap = (n * P.mean_ins_size - int(res.obs_sum[e])) / float(n)
            sd0 = 2 ** 32
        assert int(res.gap[e]) == int(gap)
        if -gap > len1 or -gap > len2:
            assert res.score[e] == 0 and res.flags[e] & abi.EDGE_NEGGAP
            continue
        assert not res.flags[e] & abi.EDGE_NEGGAP
        sd = float(res.sd_obs[e])
        sds = 0 if (sd == 0 or sd != sd) else min(sd / sd0, sd0 / sd)
        span = 0 if n < 5 else 1 - float(res.ks[e])
        want = sds + span if sds > 0.5 and span > 0.5 else 0
        assert res.score[e] == want
        changed += int(res.gap[e] != base_gap[e])
    assert changed > 20   # the lognormal estimator really replaced the normal one
