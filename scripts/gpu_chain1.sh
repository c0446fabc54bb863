#!/bin/bash
# 1 GPU: full parity suite, then decode benches: default vs no fusion vs no chaining
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_chain.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_chain.log
for m in smollm-135m:q8_0:0 smollm-1.7b:q4_0:512; do
  IFS=: read model kind ctx <<< "$m"
  timeout 300 python scripts/bench_decode.py --model $model --kind $kind --tokens 64 --context $ctx > gpurun_out/decode_${model}_chain.log 2>&1
  ZG_CUDA_FUSE=0 timeout 300 python scripts/bench_decode.py --model $model --kind $kind --tokens 64 --context $ctx > gpurun_out/decode_${model}_nofuse.log 2>&1
  ZG_CUDA_CHAIN=0 timeout 300 python scripts/bench_decode.py --model $model --kind $kind --tokens 64 --context $ctx > gpurun_out/decode_${model}_nochain.log 2>&1
done
timeout 300 python scripts/bench_sharded.py --model llama3-70b --layers 16 --emulate-world 8 --tokens 32 --batch 1,8 --context 512 > gpurun_out/shard8_emul.log 2>&1
ZG_CUDA_FUSE=0 timeout 300 python scripts/bench_sharded.py --model llama3-70b --layers 16 --emulate-world 8 --tokens 32 --batch 1 --context 512 > gpurun_out/shard8_emul_nofuse.log 2>&1
tail -n 3 gpurun_out/pytest_chain.log; for f in gpurun_out/decode_smollm*chain.log gpurun_out/decode_smollm*nofuse.log gpurun_out/shard8_emul*.log; do echo $f; grep -h '^{' $f | cut -c1-420; done
