#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
for m in smollm-135m:q8_0:0 smollm-1.7b:q4_0:512; do
  IFS=: read model kind ctx <<< "$m"
  timeout 300 python scripts/bench_decode.py --model $model --kind $kind --tokens 128 --context $ctx > gpurun_out/decode_${model}.log 2>&1; grep -h '^{' gpurun_out/decode_${model}.log | cut -c1-260
done
python -c "import __graft_entry__ as g; g.smoke()"
