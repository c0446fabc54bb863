#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu --no-extras > gpurun_out/bench_a.log 2> gpurun_out/bench_a.err
python scripts/show_bench.py gpurun_out/bench_a.log
timeout 600 python bench.py --no-cpu --no-extras > gpurun_out/bench_b.log 2> gpurun_out/bench_b.err
python scripts/show_bench.py gpurun_out/bench_b.log
