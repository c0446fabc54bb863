echo "== old kernel"
export ZG_GEMV_STREAM=0
timeout 200 python scripts/gemv_case.py 4096 14336 --copies 8 --chain
timeout 200 python scripts/gemv_case.py 4096 4096 --copies 16 --chain
timeout 200 python scripts/gemv_case.py 8192 8192 --copies 16 --chain
for wd in 1 2 8; do EMULATE_WORLD=$wd LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90; done
timeout 300 python scripts/bench_decode.py --model llama3-8b --kind q4_0 --context 512 --tokens 32 --layers 16 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('8B/16L', d['device_tok_s'], d['value'])"
timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('1.7B', d['device_tok_s'], d['value'])"
export ZG_GEMV_STREAM=1
for mq in 1024 2048 4096 8192; do
echo "== auto MIN=$mq"
export ZG_GEMV_STREAM_MIN=$mq
for wd in 1 2 8; do EMULATE_WORLD=$wd LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90; done
timeout 300 python scripts/bench_decode.py --model llama3-8b --kind q4_0 --context 512 --tokens 32 --layers 16 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('8B/16L', d['device_tok_s'], d['value'])"
timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('1.7B', d['device_tok_s'], d['value'])"
done
