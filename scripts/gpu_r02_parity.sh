# Round 2 parity run (1 GPU): the whole -m gpu suite, then the fused decode kernel under the same llama tests.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -15 gpurun_out/r02_pytest_gpu.log
ZG_CUDA_DECODE=1 timeout 300 python -m pytest tests/test_gpu_llama.py tests/test_gpu_synth_model.py -m gpu -x -q -k "llama or smollm" > gpurun_out/r02_pytest_fused.log 2>&1; echo "fused rc=$?"; tail -5 gpurun_out/r02_pytest_fused.log
