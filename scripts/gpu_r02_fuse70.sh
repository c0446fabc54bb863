export EMULATE_WORLD=1 LAYERS=8 ZG_GEMV_STREAM=0
for f in 0 1 2 3; do
ZG_CUDA_GEMV_FUSE=$f timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-100 | sed "s/^/FUSE=$f /"
done
export ZG_GEMV_STREAM=1 ZG_GEMV_STREAM_MIN=32
for f in 0 1; do
ZG_CUDA_GEMV_FUSE=$f ZG_CUDA_GEMV_PAIR=0 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-100 | sed "s/^/STREAM32 FUSE=$f /"
done
ZG_CUDA_GEMV_PAIR=0 timeout 300 python scripts/trace_decode.py --model llama3-70b --kind q4_0 --context 512 --layers 4 --emulate-world 1 --show 20 2>&1 | tail -28
