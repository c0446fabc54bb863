#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_qmatmul.py tests/test_gpu_llama.py tests/test_gpu_conformance.py -x -q -m gpu > gpurun_out/pytest_gemm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gemm.log
timeout 300 python scripts/bench_prefill.py --kind q8_0 > gpurun_out/prefill_q8.log 2>&1
timeout 300 python scripts/bench_prefill.py --kind q4_0 > gpurun_out/prefill_q4.log 2>&1
ZG_GEMM_X1=1 timeout 300 python scripts/bench_prefill.py --kind q8_0 > gpurun_out/prefill_x1.log 2>&1
tail -n 12 gpurun_out/pytest_gemm.log; cat gpurun_out/prefill_q8.log gpurun_out/prefill_q4.log gpurun_out/prefill_x1.log
