set -x
export EMULATE_WORLD=1 LAYERS=8
timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1
for P in 1 2 4 6; do ZG_GEMV_P=$P timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1; done
for NS in 3 4; do ZG_GEMV_NS=$NS timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1; done
ZG_GEMV_WAVE=1 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1
ZG_CUDA_GEMV_PAIR=0 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1
timeout 300 python scripts/trace_decode.py --model llama3-70b --kind q4_0 --context 512 --layers 4 --emulate-world 1 --show 30 2>&1 | tail -45
