#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
for m in smollm-135m:q8_0:0 smollm-1.7b:q4_0:512; do
  IFS=: read model kind ctx <<< "$m"
  timeout 300 python scripts/bench_decode.py --model $model --kind $kind --tokens 128 --context $ctx --cpu-tokens 2 > gpurun_out/decode_${model}.log 2>&1; grep -h '^{' gpurun_out/decode_${model}.log | cut -c1-250; grep -ho '"greedy_tokens_match_cpu": [a-z]*' gpurun_out/decode_${model}.log
  ZG_CUDA_GEMV_FUSE=0 timeout 300 python scripts/bench_decode.py --model $model --kind $kind --tokens 128 --context $ctx > gpurun_out/decode_${model}_nogf.log 2>&1; grep -h '^{' gpurun_out/decode_${model}_nogf.log | cut -c1-250
done
timeout 300 python scripts/bench_sharded.py --model llama3-70b --layers 16 --emulate-world 8 --tokens 32 --batch 1 --context 512 > gpurun_out/shard8_emul.log 2>&1; grep -h '^{' gpurun_out/shard8_emul.log | cut -c90-250
