#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
for m in smollm-135m:q8_0:0 smollm-1.7b:q4_0:512; do
  IFS=: read model kind ctx <<< "$m"
  timeout 300 python scripts/bench_decode.py --model $model --kind $kind --tokens 128 --context $ctx --cpu-tokens 3 > gpurun_out/decode_${model}.log 2>&1
done
timeout 300 python scripts/bench_prefill.py --kind q8_0 > gpurun_out/prefill_q8.log 2>&1
timeout 300 python scripts/trace_decode.py --model smollm-1.7b --show 30 > gpurun_out/trace_1p7b.log 2>&1
timeout 300 python scripts/trace_decode.py --model llama3-70b --layers 8 --emulate-world 8 --show 30 > gpurun_out/trace_70b_w8.log 2>&1
bash scripts/gpu_profile_all.sh > gpurun_out/profile_all.log 2>&1
python scripts/show_bench.py gpurun_out/bench_default.log
for f in gpurun_out/decode_smollm-135m.log gpurun_out/decode_smollm-1.7b.log gpurun_out/prefill_q8.log; do grep -h '^{' $f | cut -c1-420; done
