"""Host-side scalar helpers the hot path needs once per library:
`MaxObsDistr` (mathstats.normaldist.normal; call sites libmetrics.py:23,
CreateGraph.py:952,966).  Restated from the published formulas (A&S 26.2.23);
the package itself is not vendored by the reference -- PARITY UNPINNED
(SURVEY.md 8c)."""
from math import log, sqrt


def _rational(t):
    num = (0.010328 * t + 0.802853) * t + 2.515517
    den = ((0.001308 * t + 0.189269) * t + 1.432788) * t + 1.0
    return t - num / den


def normal_cdf_inverse(p):
    if not 0.0 < p < 1.0:
        raise ValueError("p must be in (0, 1)")
    if p < 0.5:
        return -_rational(sqrt(-2.0 * log(p)))
    return _rational(sqrt(-2.0 * log(1.0 - p)))


def MaxObsDistr(nr_of_obs, prob):
    """k such that P(N(0,1) < k) = prob ** (1/n)."""
    p = 1 - prob ** (1 / float(nr_of_obs))
    return normal_cdf_inverse(1 - p)
