#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 2 gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_sharded.py --model llama3-70b --layers 8 --emulate-world 8 --tokens 16 --batch 1,8 --context 512 > gpurun_out/rows_w8.log 2>&1
timeout 300 python scripts/bench_sharded.py --model llama3-70b --layers 4 --emulate-world 2 --tokens 16 --batch 1,8 --context 512 > gpurun_out/rows_w2.log 2>&1
grep -h '^{' gpurun_out/rows_w8.log gpurun_out/rows_w2.log | cut -c90-260
timeout 300 python scripts/trace_decode.py --model llama3-70b --layers 4 --emulate-world 8 --batch 8 --show 30 2>&1 | tail -42
