set -x
mkdir -p gpurun_out
ZG_CUDA_DECODE=1 timeout 200 python -m pytest tests/test_gpu_llama.py -x -q 2>&1 | tail -3
ZG_CUDA_DECODE=1 timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 --cpu-tokens 1 > gpurun_out/r02_decode_1p7b_fused.json 2>gpurun_out/err.txt; cat gpurun_out/r02_decode_1p7b_fused.json; tail -3 gpurun_out/err.txt
ZG_CUDA_DECODE=1 timeout 300 python scripts/bench_decode.py --model smollm-135m --kind q8_0 --context 0 --tokens 64 --cpu-tokens 1 > gpurun_out/r02_decode_135m_fused.json 2>gpurun_out/err.txt; cat gpurun_out/r02_decode_135m_fused.json; tail -3 gpurun_out/err.txt
ZG_CUDA_DECODE=1 timeout 200 python scripts/trace_decode.py --model smollm-1.7b --kind q4_0 --context 512 --layers 6 --show 5 > gpurun_out/r02_trace_1p7b_fused.txt 2>&1; head -64 gpurun_out/r02_trace_1p7b_fused.txt
