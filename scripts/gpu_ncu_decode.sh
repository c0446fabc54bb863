#!/bin/bash
# ncu launch list (per-kernel durations, serialized) of two SmolLM-1.7B Q4_0 decode steps at context 512
mkdir -p gpurun_out
for mode in chain nochain; do
  if [ $mode = nochain ]; then export ZG_CUDA_CHAIN=0; else unset ZG_CUDA_CHAIN; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1200 -c 700 --csv --log-file gpurun_out/decode_launches_$mode.csv \
    python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --tokens 3 --context 512 > gpurun_out/ncu_decode_$mode.log 2>&1
done
wc -l gpurun_out/decode_launches_*.csv
