mkdir -p gpurun_out
ZG_GEMV_STREAM=2 timeout 900 python -m pytest tests/test_gpu_qmatmul.py tests/test_gpu_llama.py tests/test_gpu_synth_model.py -m gpu -x -q 2>&1 | tail -15
for mq in 0 4 8 16 32; do
export ZG_GEMV_STREAM_MIN=$mq
if [ $mq = 0 ]; then export ZG_GEMV_STREAM=0; else export ZG_GEMV_STREAM=1; fi
for wd in 1 8; do
EMULATE_WORLD=$wd LAYERS=8 ZG_CUDA_GEMV_PAIR=0 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90 | sed "s/^/MIN=$mq nopair /"
done
EMULATE_WORLD=1 LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90 | sed "s/^/MIN=$mq pair /"
timeout 300 python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --context 512 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('MIN=$mq 1.7B', d['device_tok_s'], d['value'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --no-check --decode-layers 2 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('MIN=$mq gemv', d['gemv']['value'], [c['gbps'] for c in d['gemv']['cases']])"
done
