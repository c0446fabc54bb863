#!/bin/bash
mkdir -p gpurun_out
for f in 0 1 2; do
  echo "GEMV_FUSE=$f"
  ZG_CUDA_GEMV_FUSE=$f timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --tokens 128 --context 512 2>&1 | grep -h '^{' | grep -o '"device_ms_per_token": [0-9.]*\|"kernels_per_token": [0-9]*' | paste - -
  ZG_CUDA_GEMV_FUSE=$f timeout 300 python scripts/bench_sharded.py --model llama3-70b --layers 16 --emulate-world 8 --tokens 32 --batch 1 --context 512 2>&1 | grep -h '^{' | grep -o '"device_ms_per_step": [0-9.]*\|"kernels_per_step": [0-9]*' | paste - -
done
ZG_CUDA_GEMV_FUSE=1 timeout 300 python scripts/trace_decode.py --model smollm-1.7b --show 26 2>&1 | tail -36
