"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel
totals and shares (cold-cache, serialised: compare SHARES, not absolutes)."""
import collections
import csv
import sys


def main(path, ours_only=True):
    """ours_only: keep this library's kernels (k_*); the torch kernels of the synthetic generator and of
    bench.py's bookkeeping run outside the timed region."""
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].replace("void <unnamed>::", "").replace("<unnamed>::", "").split("(")[0]
        if ours_only and not k.startswith("k_"):
            continue
        v = float(r["Metric Value"])
        if r.get("Metric Unit") in ("ns", "nsecond"):
            v /= 1e3
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("%d launches, %.1f us total" % (sum(len(v) for v in agg.values()), tot))
    print("%-62s %5s %11s %10s %7s" % ("kernel", "n", "sum us", "avg us", "share"))
    for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
        print("%-62s %5d %11.1f %10.1f %6.1f%%" % (k, len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))


if __name__ == "__main__":
    main(sys.argv[1], ours_only="--all" not in sys.argv)
