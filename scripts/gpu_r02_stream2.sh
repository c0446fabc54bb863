export ZG_CUDA_GEMV_PAIR=0
for early in 1 0; do
for mq in 8 16 32; do
export ZG_GEMV_STREAM_MIN=$mq ZG_GEMV_STREAM_EARLY=$early
for wd in 1 8; do
EMULATE_WORLD=$wd LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90 | sed "s/^/EARLY=$early MIN=$mq /"
done
done
done
export ZG_GEMV_STREAM_EARLY=0 ZG_GEMV_STREAM_MIN=16
timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('1.7B', d['device_tok_s'], d['value'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --no-check --decode-layers 2 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('gemv', d['gemv']['value'], [c['gbps'] for c in d['gemv']['cases']])"
timeout 300 python scripts/trace_decode.py --model llama3-70b --kind q4_0 --context 512 --layers 4 --emulate-world 1 --show 20 2>&1 | tail -22
