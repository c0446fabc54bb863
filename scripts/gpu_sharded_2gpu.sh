#!/bin/bash
# 2-GPU run: sharded parity tests (peer-memory all-reduce and NCCL), then sharded decode, peer vs NCCL
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/pytest_sharded.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_sharded.log
ZG_CUDA_PEER=0 timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -k two_gpu > gpurun_out/pytest_sharded_nccl.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_sharded_nccl.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR scripts/bench_sharded.py --model llama3-70b --layers 16 --kind q4_0 --tokens 32 --batch 1,8 --context 512 > gpurun_out/sharded_70b_l16_g2_peer.log 2>&1
ZG_CUDA_PEER=0 timeout 400 $TR scripts/bench_sharded.py --model llama3-70b --layers 16 --kind q4_0 --tokens 32 --batch 1,8 --context 512 > gpurun_out/sharded_70b_l16_g2_nccl.log 2>&1
tail -n 3 gpurun_out/pytest_sharded.log gpurun_out/pytest_sharded_nccl.log; grep -h '^{' gpurun_out/sharded_*_g2_*.log | cut -c1-420
