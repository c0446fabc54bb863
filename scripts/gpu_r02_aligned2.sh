export ZG_GEMV_STREAM=2
for al in 0 1 2; do
export ZG_GEMV_STREAM_ALIGN=$al
timeout 200 python scripts/gemv_case.py 8192 28672 --copies 8 | sed "s/^/ALIGN=$al /"
timeout 200 python scripts/gemv_case.py 4096 4096 --copies 64 | sed "s/^/ALIGN=$al /"
timeout 200 python scripts/gemv_case.py 4096 14336 --copies 32 | sed "s/^/ALIGN=$al /"
timeout 200 python scripts/gemv_case.py 8192 28672 --copies 8 --chain | sed "s/^/ALIGN=$al /"
timeout 200 python scripts/gemv_case.py 8192 28672 --copies 8 --kind q8_0 | sed "s/^/ALIGN=$al /"
done
export ZG_GEMV_STREAM=1 ZG_GEMV_STREAM_ALIGN=0
for mq in 1024 2048 4096; do
export ZG_GEMV_STREAM_MIN=$mq
for wd in 1 2 8; do EMULATE_WORLD=$wd LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90 | sed "s/^/MIN=$mq /"; done
timeout 300 python scripts/bench_decode.py --model llama3-8b --kind q4_0 --context 512 --tokens 32 --layers 16 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('MIN=$mq 8B/16L', d['device_tok_s'], d['value'])"
timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('MIN=$mq 1.7B', d['device_tok_s'], d['value'])"
done
export ZG_GEMV_STREAM=0
for wd in 1 2 8; do EMULATE_WORLD=$wd LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90 | sed "s/^/OLD /"; done
timeout 300 python scripts/bench_decode.py --model llama3-8b --kind q4_0 --context 512 --tokens 32 --layers 16 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('OLD 8B/16L', d['device_tok_s'], d['value'])"
