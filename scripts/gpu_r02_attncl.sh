timeout 900 python -m pytest tests/test_gpu_llama.py tests/test_gpu_synth_model.py tests/test_gpu_quantized_kv_program.py tests/test_gpu_dense_head.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
export ZG_CUDA_ATTN_CLUSTER=$v
for wd in 8 1; do EMULATE_WORLD=$wd LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90 | sed "s/^/ATTN_CL=$v /"; done
timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('ATTN_CL=$v 1.7B', d['device_tok_s'], d['value'])"
timeout 300 python scripts/bench_decode.py --model smollm-135m --kind q8_0 --context 512 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('ATTN_CL=$v 135M ctx512', d['device_tok_s'], d['value'])"
done
