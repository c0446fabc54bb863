# Round 2: the fused persistent decode kernel (csrc/decode.cu) — parity first, then tok/s against the general schedule.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_llama.py -x -q > gpurun_out/r02_pytest_llama.log 2>&1; echo "llama tests rc=$?"; tail -15 gpurun_out/r02_pytest_llama.log
for m in "smollm-135m q8_0 0" "smollm-1.7b q4_0 512"; do
  set -- $m
  timeout 300 python scripts/bench_decode.py --model $1 --kind $2 --context $3 --tokens 64 --cpu-tokens 2 > gpurun_out/r02_decode_$1_fused.json 2> gpurun_out/r02_decode_$1_fused.err; echo "rc=$?"; cat gpurun_out/r02_decode_$1_fused.json; tail -3 gpurun_out/r02_decode_$1_fused.err
  ZG_CUDA_DECODE=0 timeout 300 python scripts/bench_decode.py --model $1 --kind $2 --context $3 --tokens 64 > gpurun_out/r02_decode_$1_general.json 2>&1; cat gpurun_out/r02_decode_$1_general.json
done
timeout 200 python scripts/trace_decode.py --model smollm-1.7b --kind q4_0 --context 512 --show 60 > gpurun_out/r02_trace_1p7b_fused.txt 2>&1; cat gpurun_out/r02_trace_1p7b_fused.txt
nvidia-smi --query-gpu=clocks.sm,clocks_throttle_reasons.active --format=csv,noheader
