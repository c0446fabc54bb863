set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for f in 0 1; do
ZG_CUDA_ATTN_LAYER=$f timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 --cpu-tokens 1 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('1.7B attn_layer=$f', d['device_tok_s'], d['value'], d['kernels_per_token'], d.get('greedy_tokens_match_cpu'))"
ZG_CUDA_ATTN_LAYER=$f timeout 300 python scripts/bench_decode.py --model smollm-135m --kind q8_0 --context 0 --tokens 64 --cpu-tokens 2 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('135M attn_layer=$f', d['device_tok_s'], d['value'], d['kernels_per_token'], d.get('greedy_tokens_match_cpu'))"
EMULATE_WORLD=8 LAYERS=8 ZG_CUDA_ATTN_LAYER=$f timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1
EMULATE_WORLD=1 LAYERS=8 ZG_CUDA_ATTN_LAYER=$f timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1
done
