#!/usr/bin/env python
"""Print the per-case table of bench.py JSON lines found in the given log files."""
import json, sys
for f in sys.argv[1:]:
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line)
            print(f, "value", d["value"], "e2e", d.get("e2e", {}).get("value"), "frac", d["roofline"]["frac"], "clk", d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
            for c in d["cases"]:
                print("   %5d x %5d %-7s %8.3f us %8.1f GB/s  %.3f" % (c["K"], c["N"], c["format"], c["us_per_gemv"], c["gbps"], c["frac_of_measured_hbm"]))
        elif line.strip():
            print(f, line[:300].rstrip())
