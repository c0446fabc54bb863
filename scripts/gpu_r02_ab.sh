set -x
mkdir -p gpurun_out
for i in 1 2; do
for w in 1 0; do
ZG_GEMV_WAVE=$w timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras --no-check --gemv-steps 2 --decode-layers 24 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('WAVE=$w', d['value'], d['ms_per_step'])"
done
done
for w in 1 0; do
ZG_GEMV_WAVE=$w timeout 300 python scripts/bench_sharded_emulate.py 2>/dev/null
done
