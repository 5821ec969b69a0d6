"""Expected number of links spanning two contigs (host scalar, once per
library).  Same interface and arithmetic as the reference's
BESST/e_nr_links.py:18-93 (`Param`, `ExpectedLinks`): normcdf through libm
erfc, normpdf evaluated with 100-digit decimals and rounded once."""
import math
from decimal import Decimal, getcontext


def normcdf(x, mu, sigma):
    y = 0.5 * math.erfc(-(x - mu) / (sigma * math.sqrt(2.0)))
    return 1.0 if y > 1.0 else y


def normpdf(x, mu, sigma):
    getcontext().prec = 100
    u = Decimal(str(x - mu)) / Decimal(str(abs(sigma)))
    scale = 1 / Decimal(str(math.sqrt(2 * math.pi) * abs(sigma)))
    return float(str(scale * Decimal(str(-u * u / 2)).exp()))


class Param(object):
    def __init__(self, mean, stddev, cov, read_len, softclipped):
        self.mean = mean
        self.stddev = stddev
        self.read_len = read_len
        self.cov = cov
        self.softclipped = softclipped
        self.readfrequency = 2 * self.read_len / self.cov


def ExpectedLinks(len1, len2, d, param):
    sd = float(param.stddev)
    gap = max(d, 0)
    inner = param.read_len - param.softclipped
    lo, hi = min(len1, len2), max(len1, len2)
    b1 = (len1 + len2 + gap + 2 * param.softclipped - param.mean) / sd
    a1 = (hi + gap + inner - param.mean) / sd
    b2 = (lo + gap + inner - param.mean) / sd
    a2 = (gap + 2 * inner - param.mean) / sd
    f = param.readfrequency

    def part(a, b):
        e1 = (lo - inner) / f * normcdf(a, 0, 1)
        e2 = -(-param.softclipped) / f * normcdf(b, 0, 1)
        e3 = (b * sd) / f * (normcdf(b, 0, 1) - normcdf(a, 0, 1))
        e4 = (sd / f) * (normpdf(b, 0, 1) - normpdf(a, 0, 1))
        return e1 + e2 + e3 + e4

    return part(a1, b1) - part(a2, b2)
