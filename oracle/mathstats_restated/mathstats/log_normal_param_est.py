"""ORACLE / TEST INFRASTRUCTURE -- restatement of `mathstats.log_normal_param_est.GapEstimator`, the
lognormal counterpart of GapEst that the reference calls for libraries flagged `param.lognormal`
(libmetrics.py:360-390) at CreateGraph.py:526, MakeScaffolds.py:425-426 and order_contigs.py:304-306.

PARITY UNPINNED, and here doubly so: the package is not available in this container AND the reference's own
lognormal scoring branch cannot run under Python 3 (`range` with a float step, CreateGraph.py:490), so there
is no reference output to compare with.  What is restated is the MODEL of the GapEst paper (besst.bib:48-80)
with a lognormal fragment-length density, checked against numerical quadrature and against simulated
libraries with a known gap in tests/test_oracle_math.py:

  fragment length x ~ LogNormal(mu, sigma); contigs c_min <= c_max; gap d; read length r; an observation is
  o = x - d; the number of placements of a fragment with observation o is the same trapezoid as in the normal
  model,  w(o) = o-2r+1 on [2r-1, c_min+r],  c_min-r+1 on [c_min+r, c_max+r],  c_min+c_max+1-o on
  [c_max+r, c_min+c_max+1];  p(o | d) = w(o) f(o+d) / g(d),  g(d) = int w(x-d) f(x) dx  (closed form: the
  partial moments of the lognormal,  int f = Phi(z),  int x f = exp(mu + sigma^2/2) Phi(z - sigma),
  z = (ln x - mu)/sigma).

  The ML gap maximises  L(d) = sum_i [ -ln(o_i+d) - (ln(o_i+d) - mu)^2 / (2 sigma^2) ] - n ln g(d)  over the
  integers d in [max(-c_min, 1 - min o_i), int(E[x] + 4 sd[x])] (integer ternary search, L is unimodal there).

The sum over the samples is taken in a FIXED order (32 strided partial sums combined by a butterfly) so that
the C and the CUDA restatements (one warp per edge, one lane per partial sum) reproduce it bit for bit.
"""
from math import erfc, exp, log, sqrt


def _Phi(z):
    return 0.5 * erfc(-z / sqrt(2.0))


def _F0(x, mu, sigma):
    return _Phi((log(x) - mu) / sigma) if x > 0 else 0.0


def _F1(x, mu, sigma):
    return exp(mu + sigma * sigma / 2.0) * _Phi((log(x) - mu - sigma * sigma) / sigma) if x > 0 else 0.0


def g_of_d(d, mu, sigma, c_min, c_max, r):
    """int w(x - d) f(x) dx in closed form"""
    A = d + 2 * r - 1
    B = d + c_min + r
    C = d + c_max + r
    D = d + c_min + c_max + 1
    f0A, f0B, f0C, f0D = _F0(A, mu, sigma), _F0(B, mu, sigma), _F0(C, mu, sigma), _F0(D, mu, sigma)
    f1A, f1B, f1C, f1D = _F1(A, mu, sigma), _F1(B, mu, sigma), _F1(C, mu, sigma), _F1(D, mu, sigma)
    piece1 = -(d + 2 * r - 1) * (f0B - f0A) + (f1B - f1A)
    piece2 = (c_min - r + 1) * (f0C - f0B)
    piece3 = (d + c_min + c_max + 1) * (f0D - f0C) - (f1D - f1C)
    return piece1 + piece2 + piece3


def _butterfly_sum(partial):
    p = list(partial)
    off = 16
    while off > 0:
        p = [p[l] + p[l ^ off] for l in range(32)]
        off >>= 1
    return p[0]


def log_likelihood(d, mu, sigma, samples, c_min, c_max, r):
    g = g_of_d(d, mu, sigma, c_min, c_max, r)
    if not (g > 0):
        return float("-inf")
    partial = [0.0] * 32
    v2 = 2.0 * sigma * sigma
    for k, o in enumerate(samples):
        lx = log(o + d)
        partial[k & 31] += -lx - (lx - mu) * (lx - mu) / v2
    return _butterfly_sum(partial) - len(samples) * log(g)


def search_bounds(mu, sigma, samples, c_min):
    mean_x = exp(mu + sigma * sigma / 2.0)
    sd_x = sqrt((exp(sigma * sigma) - 1.0) * exp(2.0 * mu + sigma * sigma))
    d_lower = max(int(-c_min), 1 - min(samples))
    d_upper = int(mean_x + 4.0 * sd_x)
    return d_lower, d_upper


def GapEstimator(mu, sigma, read_length, samples, c1_len, c2_len=None):
    if c2_len is None:
        c2_len = c1_len
    c_min, c_max = min(c1_len, c2_len), max(c1_len, c2_len)
    lo, hi = search_bounds(mu, sigma, samples, c_min)
    if hi <= lo:
        return lo
    L = lambda d: log_likelihood(d, mu, sigma, samples, c_min, c_max, read_length)   # noqa: E731
    while hi - lo > 2:
        third = (hi - lo) // 3
        m1, m2 = lo + third, hi - third
        if L(m1) < L(m2):
            lo = m1 + 1
        else:
            hi = m2 - 1
    best, best_l = lo, L(lo)
    for d in range(lo + 1, hi + 1):
        v = L(d)
        if v > best_l:
            best, best_l = d, v
    return best
