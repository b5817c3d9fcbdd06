#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share, avg, max."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v * 1e6 if unit == "s" else v
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name)
        seq.append((name, v))
    tot, cnt, mx = collections.defaultdict(float), collections.Counter(), collections.defaultdict(float)
    for n, v in seq:
        tot[n] += v
        cnt[n] += 1
        mx[n] = max(mx[n], v)
    total = sum(tot.values())
    print(f"{len(seq)} launches, {total / 1000:.3f} ms total (cold-cache, serialised: compare SHARES)")
    for n in sorted(tot, key=lambda k: -tot[k]):
        print(f"{n:38s} n={cnt[n]:5d} total={tot[n] / 1000:9.3f} ms share={100 * tot[n] / total:5.1f}% avg={tot[n] / cnt[n]:9.1f} us max={mx[n]:9.1f} us")
    return seq


if __name__ == "__main__":
    main(sys.argv[1])
