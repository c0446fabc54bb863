#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 2 gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --tokens 128 --context 512 --cpu-tokens 3 > gpurun_out/decode_smollm-1.7b.log 2>&1; grep -h '^{' gpurun_out/decode_smollm-1.7b.log | cut -c1-250
timeout 300 python scripts/bench_decode.py --model smollm-135m --kind q8_0 --tokens 128 --context 0 --cpu-tokens 3 > gpurun_out/decode_smollm-135m.log 2>&1; grep -h '^{' gpurun_out/decode_smollm-135m.log | cut -c1-250
timeout 600 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
python scripts/show_bench.py gpurun_out/bench_default.log
