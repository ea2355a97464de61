#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel, number of
launches, mean and total device time, share of the total."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        try:
            d[r[ki].split("(")[0][-60:]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    total = sum(sum(v) for v in d.values())
    print("%-62s %6s %10s %10s %6s" % ("kernel", "n", "mean_ms", "total_ms", "share"))
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print("%-62s %6d %10.4f %10.3f %5.1f%%" % (k, len(v), sum(v) / len(v) / 1e6, sum(v) / 1e6, 100 * sum(v) / total))


if __name__ == "__main__":
    main(sys.argv[1])
