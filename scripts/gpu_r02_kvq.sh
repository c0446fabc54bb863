set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_quantized_kv_program.py tests/test_gpu_quantized_kv.py tests/test_gpu_llama.py tests/test_gpu_conformance.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python scripts/bench_kvq.py > gpurun_out/r02_kvq.json 2>gpurun_out/r02_kvq.err; tail -3 gpurun_out/r02_kvq.err; cat gpurun_out/r02_kvq.json
