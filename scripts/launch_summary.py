#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and a sample of the sequence."""
import collections
import csv
import re
import sys


def main(path, sample_from=0, sample_n=0):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = []
    for r in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        rows.append((name, float(r["Metric Value"].replace(",", "")) / 1e3, r.get("Grid Size"), r.get("Block Size")))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t, g, b in rows:
        agg[n][0] += 1
        agg[n][1] += t
    total = sum(t for _, t, _, _ in rows)
    print(f"{path}: {len(rows)} launches, {total:.1f} us (serialised, cold cache)")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {n[:60]:60s} n={c:4d} total={t:8.1f}us avg={t / c:6.2f}us share={t / total:5.1%}")
    for n, t, g, b in rows[sample_from:sample_from + sample_n]:
        print("     ", n[:44].ljust(44), f"{t:7.2f}", g, b)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
