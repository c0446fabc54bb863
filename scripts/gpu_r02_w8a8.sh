set -x
for q in 0 1; do for r in 1 2; do echo "QPDL=$q ROWS=$r"; ZG_W8A8_QPDL=$q ZG_W8A8_ROWS=$r timeout 200 python scripts/bench_w8a8.py 2>/dev/null | cut -c1-120; done; done
