for h in f32 f16 bf16; do
timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 --head $h 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('1.7B head=$h', d['device_tok_s'], d['value'], d['head_bytes_saved_per_token'], d['kernels_per_token'])"
timeout 300 python scripts/bench_decode.py --model smollm-135m --kind q8_0 --context 0 --tokens 64 --head $h 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('135M head=$h', d['device_tok_s'], d['value'], d['head_bytes_saved_per_token'], d['kernels_per_token'])"
done
