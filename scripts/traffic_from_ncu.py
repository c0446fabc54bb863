#!/usr/bin/env python
"""profiles/traffic.json from the per-case ncu captures (scripts/gpu_r02_profile.sh): DRAM bytes read + written per LAUNCH of
the matvec kernel that carries the case (the launches with the most bytes: the same capture also sees the 1-layer decode
leg of the bench command), the matvecs that launch carries (bytes / algorithmic bytes of one matvec, rounded) and the kernel."""
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
prefix = sys.argv[2] if len(sys.argv) > 2 else "r02"
BLK = {"i8_f32": 36, "q8_0": 34, "q4_0": 18}
out = {}
for path in sorted(glob.glob(os.path.join(src, f"{prefix}_traffic_*.csv"))):
    case = re.search(rf"{prefix}_traffic_(.+)\.csv", path).group(1)
    K, N, fmt = re.match(r"(\d+)x(\d+)_(.+)", case).groups()
    alg = int(K) * int(N) // 32 * BLK[fmt] + 4 * int(K) + 4 * int(N)
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if l.startswith('"')))
    per = collections.defaultdict(dict)
    for r in rows:
        per[r["ID"]][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        per[r["ID"]]["grid"] = r["Grid Size"]
        per[r["ID"]]["kernel"] = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    if not per:
        continue
    for v in per.values():
        v["bytes"] = v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]
    top = max(v["bytes"] for v in per.values())
    vals = [v for v in per.values() if v["bytes"] > 0.9 * top]
    b = int(statistics.median(v["bytes"] for v in vals))
    out[case] = {"dram_bytes_per_launch": b, "gemvs_per_launch": max(1, round(b / alg)), "algorithmic_bytes_per_gemv": alg,
                 "kernel": vals[0]["kernel"], "grid": vals[0]["grid"], "launches_captured": len(vals),
                 "ncu_us_per_launch_cold_serialised": round(statistics.median(v["gpu__time_duration.sum"] for v in vals) / 1e3, 2)}
json.dump(out, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
