for mx in 4096 1024 256; do
export ZG_CUDA_NORM_CHAIN_MAX=$mx
timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('CHAIN_MAX=$mx 1.7B', d['device_tok_s'], d['value'], d['kernels_per_token'])"
timeout 300 python scripts/bench_decode.py --model smollm-135m --kind q8_0 --context 0 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('CHAIN_MAX=$mx 135M', d['device_tok_s'], d['value'], d['kernels_per_token'])"
timeout 300 python scripts/bench_decode.py --model llama3-8b --kind q4_0 --context 512 --tokens 32 --layers 16 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('CHAIN_MAX=$mx 8B/16L', d['device_tok_s'], d['value'])"
done
ZG_CUDA_NORM_CHAIN_MAX=256 timeout 600 python -m pytest tests/test_gpu_llama.py tests/test_gpu_conformance.py -m gpu -x -q 2>&1 | tail -2
