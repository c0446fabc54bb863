set -x
timeout 200 python scripts/trace_decode.py --model smollm-1.7b --kind q4_0 --context 512 --layers 8 --show 26 2>&1 | head -50
timeout 200 python scripts/trace_decode.py --model llama3-70b --kind q4_0 --context 512 --layers 6 --emulate-world 8 --show 24 2>&1 | head -40
