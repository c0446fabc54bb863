# Multi-GPU parity: the sharded tests (self-skip above the GPU count of the box).
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02_pytest_sharded.log 2>&1; echo "sharded rc=$?"; tail -15 gpurun_out/r02_pytest_sharded.log
