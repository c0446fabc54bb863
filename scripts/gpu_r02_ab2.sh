set -x
for wd in 8 2 1; do
for f in 0 1; do
EMULATE_WORLD=$wd LAYERS=8 ZG_CUDA_DECODE=$f timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1
done
done
