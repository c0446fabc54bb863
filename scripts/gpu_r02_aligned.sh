export ZG_GEMV_STREAM=2
timeout 600 python -m pytest tests/test_gpu_qgemv_stream.py -m gpu -x -q 2>&1 | tail -3
for ns in 2 4; do for c in 16 32; do
export ZG_GEMV_STREAM_NS=$ns ZG_GEMV_STREAM_CHUNKS=$c
timeout 200 python scripts/gemv_case.py 8192 28672 --copies 8 --chain | sed "s/^/NS=$ns C=$c /"
timeout 200 python scripts/gemv_case.py 8192 28672 --copies 8 | sed "s/^/NS=$ns C=$c /"
timeout 200 python scripts/gemv_case.py 4096 4096 --copies 64 | sed "s/^/NS=$ns C=$c /"
timeout 200 python scripts/gemv_case.py 4096 14336 --copies 32 | sed "s/^/NS=$ns C=$c /"
ZG_GEMV_STREAM=1 ZG_CUDA_GEMV_PAIR=0 EMULATE_WORLD=1 LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90 | sed "s/^/NS=$ns C=$c auto /"
done; done
