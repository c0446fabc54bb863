for cfg in "1 128" "2 128" "2 64" "4 64" "4 32"; do
set -- $cfg
export ZG_CUDA_ATTN_SPLIT_MULT=$1 ZG_CUDA_ATTN_MIN_POS=$2
for wd in 1 8; do EMULATE_WORLD=$wd LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90 | sed "s/^/MULT=$1 MINPOS=$2 /"; done
timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('MULT=$1 MINPOS=$2 1.7B', d['device_tok_s'], d['value'])"
done
