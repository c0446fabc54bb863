set -x
mkdir -p gpurun_out
timeout 200 python scripts/trace_decode.py --model smollm-1.7b --kind q4_0 --context 512 --layers 6 --show 70 > gpurun_out/r02_trace_1p7b_fused.txt 2>&1; cat gpurun_out/r02_trace_1p7b_fused.txt
