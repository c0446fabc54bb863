#!/bin/bash
# 1 GPU: rank 0's shard of an 8-way Llama-3-70B split (collectives = identity): timing, then ncu launch list
mkdir -p gpurun_out
timeout 300 python scripts/bench_sharded.py --model llama3-70b --layers 16 --emulate-world 8 --tokens 32 --batch 1,8 --context 512 > gpurun_out/shard8_emul.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 -c 400 --csv --log-file gpurun_out/shard8_launches_b1.csv \
    python scripts/bench_sharded.py --model llama3-70b --layers 8 --emulate-world 8 --tokens 3 --batch 1 --context 512 > gpurun_out/ncu_shard8_b1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 800 -c 400 --csv --log-file gpurun_out/shard8_launches_b8.csv \
    python scripts/bench_sharded.py --model llama3-70b --layers 8 --emulate-world 8 --tokens 3 --batch 8 --context 512 > gpurun_out/ncu_shard8_b8.log 2>&1
grep -h '^{' gpurun_out/shard8_emul.log
