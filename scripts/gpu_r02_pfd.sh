for f in 0 2 4 8; do
export ZG_GEMV_PFD=$f
for wd in 1 8; do
EMULATE_WORLD=$wd LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90 | sed "s/^/PFD=$f /"
done
timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('PFD=$f 1.7B', d['device_tok_s'], d['value'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --no-check --decode-layers 2 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('PFD=$f gemv', d['gemv']['value'], [c['gbps'] for c in d['gemv']['cases']])"
done
