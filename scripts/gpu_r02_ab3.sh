set -x
for wd in 8 4 1; do
for f in 0 1; do
EMULATE_WORLD=$wd LAYERS=8 ZG_CUDA_L2_PREFETCH=$f timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1
done
done
for f in 0 1; do
ZG_CUDA_L2_PREFETCH=$f timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('1.7B prefetch=$f', d['device_tok_s'], d['value'])"
ZG_CUDA_L2_PREFETCH=$f timeout 300 python scripts/bench_decode.py --model smollm-135m --kind q8_0 --context 0 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('135M prefetch=$f', d['device_tok_s'], d['value'])"
ZG_CUDA_L2_PREFETCH=$f timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --no-check --decode-layers 2 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('gemv prefetch=$f', d['gemv']['value'], [c['gbps'] for c in d['gemv']['cases']])"
done
