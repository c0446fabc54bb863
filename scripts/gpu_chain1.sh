#!/bin/bash
# 1 GPU: full parity suite with chained small ops, then decode benches with / without chaining
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_chain.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_chain.log
for m in smollm-135m:q8_0:0 smollm-1.7b:q4_0:512; do
  IFS=: read model kind ctx <<< "$m"
  timeout 300 python scripts/bench_decode.py --model $model --kind $kind --tokens 64 --context $ctx --profile > gpurun_out/decode_${model}_chain.log 2>&1
  ZG_CUDA_CHAIN=0 timeout 300 python scripts/bench_decode.py --model $model --kind $kind --tokens 64 --context $ctx > gpurun_out/decode_${model}_nochain.log 2>&1
done
timeout 300 python scripts/bench_decode.py --model llama3-70b --layers 4 --kind q4_0 --tokens 32 --context 512 --profile > gpurun_out/decode_70b_l4_chain.log 2>&1
ZG_CUDA_CHAIN=0 timeout 300 python scripts/bench_decode.py --model llama3-70b --layers 4 --kind q4_0 --tokens 32 --context 512 > gpurun_out/decode_70b_l4_nochain.log 2>&1
tail -n 3 gpurun_out/pytest_chain.log; for f in gpurun_out/decode_*chain.log; do echo $f; tail -n 1 $f; done
