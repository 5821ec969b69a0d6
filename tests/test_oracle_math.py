"""Pins for the oracle's restated arithmetic (mathstats is not vendored by the
reference: PARITY UNPINNED at that boundary).  What can be pinned:
the closed forms against numerical quadrature of the published GapEst model, the
C restatement against the Python restatement bit for bit, KS against scipy, the
inverse normal CDF against scipy within its published error, and e_nr_links
against the reference module's value recorded in SURVEY.md 8c."""
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
