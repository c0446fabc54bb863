set -x
mkdir -p gpurun_out
timeout 300 python scripts/trace_decode.py --model llama3-70b --kind q4_0 --context 512 --layers 6 --emulate-world 8 --show 64 > gpurun_out/r02_trace_70b_shard8.txt 2>&1; cat gpurun_out/r02_trace_70b_shard8.txt
timeout 300 python scripts/trace_decode.py --model llama3-70b --kind q4_0 --context 512 --layers 4 --emulate-world 1 --show 40 > gpurun_out/r02_trace_70b_1gpu.txt 2>&1; head -60 gpurun_out/r02_trace_70b_1gpu.txt
