"""ORACLE / TEST INFRASTRUCTURE -- restatement of `mathstats.normaldist.normal`.

`mathstats` (PyPI, pinned 0.2.6.5 by the reference: requirements.txt:3,
setup.py:33) is NOT vendored under /root/reference and is not installable here
(no network).  This file restates the published algorithms the reference
relies on at its call sites (libmetrics.py:14,23; CreateGraph.py:34,952,966):

  erf         Abramowitz & Stegun 7.1.26 (5-term polynomial, |err| <= 1.5e-7);
              docs/VERSIONS.txt:47 of the reference: "closed formula
              approximation for erf(x)".
  normpdf     plain fp64 Gaussian density.
  MaxObsDistr k with P(N(0,1) < k) = prob**(1/n), inverse CDF by A&S 26.2.23.

PARITY UNPINNED: no test of the reference pins results at this boundary
(SURVEY.md 8c); constants and operation order here are the published formulas.
"""
from math import exp, log, pi, sqrt


def erf(x):
    a1 = 0.254829592
    a2 = -0.284496736
    a3 = 1.421413741
    a4 = -1.453152027
    a5 = 1.061405429
    p = 0.3275911
    sign = 1
    if x < 0:
        sign = -1
    x = abs(x)
    t = 1.0 / (1.0 + p * x)
    y = 1.0 - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * exp(-x * x)
    return sign * y


def normcdf(x, mu, sigma):
    return 0.5 * (1.0 + erf((x - mu) / (sigma * sqrt(2.0))))


def normpdf(x, mu, sigma):
    u = (x - mu) / abs(sigma)
    return (1 / (sqrt(2 * pi) * abs(sigma))) * exp(-u * u / 2)


def _rational_approximation(t):
    # A&S 26.2.23, |err| < 4.5e-4
    c0, c1, c2 = 2.515517, 0.802853, 0.010328
    d0, d1, d2 = 1.432788, 0.189269, 0.001308
    numerator = (c2 * t + c1) * t + c0
    denominator = ((d2 * t + d1) * t + d0) * t + 1.0
    return t - numerator / denominator


def normal_CDF_inverse(p):
    assert 0.0 < p < 1.0
    if p < 0.5:
        return -_rational_approximation(sqrt(-2.0 * log(p)))
    return _rational_approximation(sqrt(-2.0 * log(1.0 - p)))


def MaxObsDistr(nr_of_obs, prob):
    p = 1 - prob ** (1 / float(nr_of_obs))
    return normal_CDF_inverse(1 - p)
