#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_chain.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_chain.log
timeout 300 python scripts/trace_decode.py --model smollm-1.7b --show 24 > gpurun_out/trace_1p7b_chain.log 2>&1
ZG_CUDA_CHAIN=0 timeout 300 python scripts/trace_decode.py --model smollm-1.7b --show 30 > gpurun_out/trace_1p7b_nochain.log 2>&1
timeout 300 python scripts/trace_decode.py --model llama3-70b --layers 8 --emulate-world 8 --show 24 > gpurun_out/trace_70b_w8_chain.log 2>&1
for m in smollm-135m:q8_0:0 smollm-1.7b:q4_0:512; do
  IFS=: read model kind ctx <<< "$m"
  timeout 300 python scripts/bench_decode.py --model $model --kind $kind --tokens 64 --context $ctx > gpurun_out/decode_${model}_chain.log 2>&1
  ZG_CUDA_CHAIN=0 timeout 300 python scripts/bench_decode.py --model $model --kind $kind --tokens 64 --context $ctx > gpurun_out/decode_${model}_nochain.log 2>&1
done
timeout 300 python scripts/bench_sharded.py --model llama3-70b --layers 16 --emulate-world 8 --tokens 32 --batch 1 --context 512 > gpurun_out/shard8_emul.log 2>&1
ZG_CUDA_CHAIN=0 timeout 300 python scripts/bench_sharded.py --model llama3-70b --layers 16 --emulate-world 8 --tokens 32 --batch 1 --context 512 > gpurun_out/shard8_emul_nochain.log 2>&1
tail -n 2 gpurun_out/pytest_chain.log
cat gpurun_out/trace_1p7b_chain.log gpurun_out/trace_70b_w8_chain.log
for f in gpurun_out/decode_smollm*chain.log gpurun_out/shard8_emul*.log; do echo $f; grep -h '^{' $f | cut -c1-330; done
