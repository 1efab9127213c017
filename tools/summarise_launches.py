#!/usr/bin/env python3
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the profiled run)."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    tot = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0]
        tot.setdefault(k, [0, 0.0])
        tot[k][0] += 1
        tot[k][1] += float(r["Metric Value"]) / 1e6
    s = sum(v[1] for v in tot.values())
    print(f"{len(rows)} launches, {s:.3f} ms total device time (ncu: cold-cache, serialised)")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:28s} n={v[0]:4d} total_ms={v[1]:10.3f} mean_ms={v[1] / v[0]:9.4f} share={v[1] / s:.4f}")


if __name__ == "__main__":
    main(sys.argv[1])
