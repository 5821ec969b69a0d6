"""ORACLE / TEST INFRASTRUCTURE -- restatement of
`mathstats.normaldist.truncatedskewed.param_est` (GapEst, Sahlin et al. 2012;
besst.bib:48-80 of the reference).  Call sites in the reference:
CreateGraph.py:537,555; MakeScaffolds.py:68,449,453; order_contigs.py:300,308;
pathgaps.py:108,204.

Model: fragment x ~ N(mu, sigma^2); contigs c_min <= c_max; gap
d; read length r; observation o = x - d; number of placements
  w(o) = o-2r+1            on [2r, c_min+r]
       = c_min-r+1         on [c_min+r, c_max+r]
       = c_min+c_max-o+1   on [c_max+r, c_min+c_max]
w is continuous and reaches zero at o = 2r-1 and o = c_min+c_max+1, so
g(d) = int w(x-d) phi(x) dx integrates over x-mu in [A, D] with
  A = d+2r-1-mu, B = c_min+d+r-mu, C = c_max+d+r-mu, D = c_min+c_max+d+1-mu
(closed form below), and the ML equation for n observations with mean obs is
  mu - mean_obs = d + sigma^2 g'(d)/g(d).

PARITY UNPINNED: the package source is not available in this container; the
closed forms below are derived from the model and checked against numerical
quadrature of that model in tests/test_oracle_math.py.  Division by an exactly-zero g(d)
follows IEEE-754 (inf/nan) instead of raising, so that the C and CUDA
restatements can agree with this file bit for bit in control flow.
"""
import sys
from math import exp, pi, sqrt, isnan

from mathstats.normaldist.normal import erf


def _div(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a == 0 or isnan(a):
            return float("nan")
        return float("inf") if a > 0 else float("-inf")


def _terms(d, mean, stdDev, c_min, c_max, readLen):
    s2 = (2 ** 0.5) * stdDev
    A = d + 2 * readLen - 1 - mean
    B = c_min + d + readLen - mean
    C = c_max + d + readLen - mean
    D = c_min + c_max + d + 1 - mean
    eA, eB, eC, eD = erf(A / s2), erf(B / s2), erf(C / s2), erf(D / s2)
    v2 = float(2 * stdDev ** 2)
    xA, xB, xC, xD = exp(-(A ** 2) / v2), exp(-(B ** 2) / v2), exp(-(C ** 2) / v2), exp(-(D ** 2) / v2)
    return eA, eB, eC, eD, xA, xB, xC, xD


def _g_and_gprime(d, mean, stdDev, c_min, c_max, readLen):
    eA, eB, eC, eD, xA, xB, xC, xD = _terms(d, mean, stdDev, c_min, c_max, readLen)
    term1 = (c_min - readLen + 1) / 2.0 * (eC - eB)
    term2 = (c_min + c_max + d - mean + 1) / 2.0 * (eD - eC)
    term3 = (d + 2 * readLen - mean - 1) / 2.0 * (eA - eB)
    k = stdDev / ((2 * pi) ** 0.5)
    term4 = k * (xD + xA)
    term5 = -k * (xC + xB)
    g_d = term1 + term2 + term3 + term4 + term5
    g_prime_d = 0.5 * (eA - eB) + 0.5 * (eD - eC)
    g_bis_d = (xA - xB - xC + xD) / (((2 * pi) ** 0.5) * stdDev)
    return g_d, g_prime_d, g_bis_d


def funcDGeneral(d, mean, stdDev, c1Len, c2Len, readLen):
    c_min = min(c1Len, c2Len)
    c_max = max(c1Len, c2Len)
    g_d, g_prime_d, _ = _g_and_gprime(d, mean, stdDev, c_min, c_max, readLen)
    Aofd = _div(g_prime_d, g_d)
    func_of_d = d + Aofd * stdDev ** 2
    return func_of_d, Aofd


def CalcMLvaluesOfdGeneral(mean, stdDev, readLen, c1Len, obs, c2Len):
    # binary search for d with funcDGeneral(d) == obs (= mean - mean_obs)
    d_upper = int(mean + 2 * stdDev - 2 * readLen)
    d_lower = int(-4 * stdDev)
    while d_upper - d_lower > 1:
        d_ML = (d_upper + d_lower) / 2.0
        func_of_d, Aofd = funcDGeneral(d_ML, mean, stdDev, c1Len, c2Len, readLen)
        if func_of_d > obs:
            d_upper = d_ML
        else:
            d_lower = d_ML
    d_ML = (d_upper + d_lower) / 2.0
    return int(round(d_ML, 0))


def GapEstimator(mean, sigma, read_length, mean_obs, c1_len, c2_len=None):
    naive_gap = mean - mean_obs
    if c2_len is None:
        c2_len = c1_len
    if c1_len <= 0 or c2_len <= 0:
        sys.stderr.write('ERROR! Gap estimation on contigs with negative length\n')
    return CalcMLvaluesOfdGeneral(mean, sigma, read_length, c1_len, naive_gap, c2_len)


def PreCalcMLvaluesOfdLongContigs(mean, stdDev, readLen):
    d_upper = int(mean + 2 * stdDev - 2 * readLen)
    d_lower = int(-4 * stdDev)
    dValuesTable = {}
    prev_obs = d_lower
    for d in range(d_lower, d_upper + 1):
        func_of_d, Aofd = funcDGeneral(d, mean, stdDev, mean + 4 * stdDev, mean + 4 * stdDev, readLen)
        obs = int(round(func_of_d, 0))
        dValuesTable[obs] = d
        if abs(obs - prev_obs) > 1:
            n = abs(obs - prev_obs)
            for i in range(0, n):
                dValuesTable[prev_obs + i + 1] = d
        prev_obs = obs
    return dValuesTable


def tr_sk_mean(mean, stdDev, readLen, c1Len, c2Len, d):
    c_min = min(c1Len, c2Len)
    c_max = max(c1Len, c2Len)
    g_d, g_prime_d, _ = _g_and_gprime(d, mean, stdDev, c_min, c_max, readLen)
    return mean - stdDev ** 2 * _div(g_prime_d, g_d) - d


def tr_sk_std_dev(mean, stdDev, readLen, c1Len, c2Len, d):
    """sqrt(E[O^2] - E[O]^2) of the observation O = x - d under the density
    w(x-d) phi(x)/g(d):  E[x] = mu - sigma^2 g'/g,
    E[x^2] = sigma^2 + mu^2 + sigma^4 g''/g - 2 mu sigma^2 g'/g."""
    c_min = min(c1Len, c2Len)
    c_max = max(c1Len, c2Len)
    g_d, g_prime_d, g_bis_d = _g_and_gprime(d, mean, stdDev, c_min, c_max, readLen)
    r1 = _div(g_prime_d, g_d)
    r2 = _div(g_bis_d, g_d)
    e_x = mean - stdDev ** 2 * r1
    e_x_square = stdDev ** 2 + mean ** 2 + stdDev ** 4 * r2 - 2 * mean * stdDev ** 2 * r1
    e_o = e_x - d
    e_o_square = e_x_square - 2 * d * e_x + d ** 2
    var = e_o_square - e_o ** 2
    if not (var >= 0):
        return 0.0
    return var ** 0.5
