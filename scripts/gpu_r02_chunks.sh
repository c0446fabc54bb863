export ZG_GEMV_STREAM=2
for c in 16 24 32 48 64 96; do
export ZG_GEMV_STREAM_CHUNKS=$c
timeout 200 python scripts/gemv_case.py 8192 28672 --copies 8 --chain | sed "s/^/C=$c /"
timeout 200 python scripts/gemv_case.py 8192 28672 --copies 8 | sed "s/^/C=$c /"
timeout 200 python scripts/gemv_case.py 4096 4096 --copies 64 | sed "s/^/C=$c /"
timeout 200 python scripts/gemv_case.py 4096 14336 --copies 32 | sed "s/^/C=$c /"
timeout 200 python scripts/gemv_case.py 8192 8192 --copies 16 --chain | sed "s/^/C=$c /"
timeout 200 python scripts/gemv_case.py 4096 14336 --copies 8 --chain | sed "s/^/C=$c /"
done
