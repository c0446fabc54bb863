mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for st in 0 1; do
export ZG_GEMV_STREAM=$st
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --no-check --decode-layers 2 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('STREAM=$st gemv', d['gemv']['value'], [c['gbps'] for c in d['gemv']['cases']], d['roofline']['frac'])"
done
for wd in 1 2 4 8; do EMULATE_WORLD=$wd LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90; done
timeout 300 python scripts/bench_decode.py --model smollm-135m --kind q8_0 --context 0 --tokens 64 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('135M', d['device_tok_s'], d['value'])"
