for ns in 2 3 4; do
export ZG_GEMV_PAIR_NS=$ns
for wd in 8 4; do EMULATE_WORLD=$wd LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90 | sed "s/^/PAIR_NS=$ns /"; done
timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('PAIR_NS=$ns 1.7B', d['device_tok_s'], d['value'])"
timeout 300 python scripts/bench_decode.py --model smollm-135m --kind q8_0 --context 0 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('PAIR_NS=$ns 135M', d['device_tok_s'], d['value'])"
done
