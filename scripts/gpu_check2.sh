#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_sharded.py --model llama3-70b --layers 16 --emulate-world 8 --tokens 32 --batch 1,8 --context 512 --trace > gpurun_out/shard8_emul.log 2> gpurun_out/shard8_emul.err
grep -h '^{' gpurun_out/shard8_emul.log | cut -c90-250; grep -A10 "^trace batch 1" gpurun_out/shard8_emul.err | head -12
timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --tokens 128 --context 512 > gpurun_out/decode_smollm-1.7b.log 2>&1; grep -h '^{' gpurun_out/decode_smollm-1.7b.log | cut -c1-260
