#!/bin/bash
# 2-GPU run: sharded parity tests, then sharded decode benches (SmolLM-1.7B, Llama-3-70B shapes)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.log 2>&1
free -g > gpurun_out/host_mem.log; nproc >> gpurun_out/host_mem.log
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/pytest_sharded.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_sharded.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/bench_sharded.py --model smollm-1.7b --kind q4_0 --tokens 64 --context 512 > gpurun_out/sharded_1p7b_g2.log 2>&1
timeout 300 python scripts/bench_sharded.py --model smollm-1.7b --kind q4_0 --tokens 64 --context 512 > gpurun_out/sharded_1p7b_g1.log 2>&1
timeout 600 $TR scripts/bench_sharded.py --model llama3-70b --kind q4_0 --tokens 16 --context 512 > gpurun_out/sharded_70b_g2.log 2>&1
tail -3 gpurun_out/pytest_sharded.log gpurun_out/sharded_*.log
