#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python bench.py --no-cpu --no-extras --steps 100 > gpurun_out/sw_$name.log 2>&1; echo "== $name"; python scripts/show_bench.py gpurun_out/sw_$name.log; }
run base A=1
run ns2 ZG_GEMV_NS=2
