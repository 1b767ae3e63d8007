#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / mean / min / max (us) and share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, rows = rows[0], rows[1:]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[ik][:80], []).append(float(r[iv].replace(",", "")) / 1000.0)
    total = sum(sum(v) for v in agg.values())
    print("%-82s %5s %9s %9s %9s %7s" % ("kernel", "n", "mean_us", "min_us", "max_us", "share"))
    for k, v in agg.items():
        print("%-82s %5d %9.2f %9.2f %9.2f %6.1f%%" % (k, len(v), sum(v) / len(v), min(v), max(v), 100 * sum(v) / total))


if __name__ == "__main__":
    main(sys.argv[1])
