#!/usr/bin/env python
"""profiles/traffic.json from the per-case ncu captures (scripts/gpu_profile_all.sh): DRAM bytes read + written per
LAUNCH of the matvec kernel (median over the captured launches of the most common grid), with the number of matvecs that
launch carries (grid.y: same-level same-shape matvecs share a launch)."""
import collections
import csv
import glob
import json
import os
import re
import statistics
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "profiles")
out = {}
for path in sorted(glob.glob(os.path.join(src, "r01_traffic_*.csv"))):
    case = re.search(r"r01_traffic_(.+)\.csv", path).group(1)
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if l.startswith('"')))
    per = collections.defaultdict(dict)
    for r in rows:
        per[r["ID"]][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        per[r["ID"]]["grid"] = r["Grid Size"]
    grids = collections.Counter(v["grid"] for v in per.values())
    if not grids:
        continue
    grid = grids.most_common(1)[0][0]
    vals = [v for v in per.values() if v["grid"] == grid]
    gy = int(re.findall(r"\d+", grid)[1])
    out[case] = {"dram_bytes_per_launch": int(statistics.median(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for v in vals)),
                 "gemvs_per_launch": gy, "grid": grid, "launches_captured": len(vals),
                 "ncu_us_per_launch_cold_serialised": round(statistics.median(v["gpu__time_duration.sum"] for v in vals) / 1e3, 2)}
json.dump(out, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
