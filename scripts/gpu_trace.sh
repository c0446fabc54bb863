#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/trace_decode.py --model smollm-1.7b --show 22 > gpurun_out/trace_1p7b.log 2>&1
timeout 300 python scripts/trace_decode.py --model llama3-70b --layers 8 --emulate-world 8 --show 22 > gpurun_out/trace_70b_w8.log 2>&1
for m in smollm-135m:q8_0:0 smollm-1.7b:q4_0:512; do
  IFS=: read model kind ctx <<< "$m"
  timeout 300 python scripts/bench_decode.py --model $model --kind $kind --tokens 128 --context $ctx > gpurun_out/decode_${model}.log 2>&1
  ZG_CUDA_ATTN_SPLIT=0 timeout 300 python scripts/bench_decode.py --model $model --kind $kind --tokens 128 --context $ctx > gpurun_out/decode_${model}_nosplit.log 2>&1
done
timeout 300 python scripts/bench_sharded.py --model llama3-70b --layers 16 --emulate-world 8 --tokens 32 --batch 1,8 --context 512 > gpurun_out/shard8_emul.log 2>&1
tail -n 2 gpurun_out/pytest_gpu.log
head -8 gpurun_out/trace_1p7b.log; tail -n 22 gpurun_out/trace_70b_w8.log
for f in gpurun_out/decode_smollm*.log gpurun_out/shard8_emul.log; do echo $f; grep -h '^{' $f | cut -c1-330; done
