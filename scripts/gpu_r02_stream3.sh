export ZG_CUDA_GEMV_PAIR=0
for ns in 2 3 4 6; do
for mq in 16 32; do
export ZG_GEMV_STREAM_MIN=$mq ZG_GEMV_STREAM_NS=$ns
EMULATE_WORLD=1 LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90 | sed "s/^/NS=$ns MIN=$mq /"
done
ZG_GEMV_STREAM=2 timeout 200 python scripts/gemv_case.py 8192 28672 --copies 8 --chain | sed "s/^/NS=$ns /"
ZG_GEMV_STREAM=2 timeout 200 python scripts/gemv_case.py 8192 28672 --copies 8 | sed "s/^/NS=$ns /"
ZG_GEMV_STREAM=2 timeout 200 python scripts/gemv_case.py 4096 4096 --copies 64 | sed "s/^/NS=$ns /"
done
