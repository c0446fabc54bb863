export ZG_GEMV_STREAM=2
for wv in 1 2 3 4 6; do
export ZG_GEMV_STREAM_WAVES=$wv ZG_GEMV_STREAM_MIN=4
timeout 200 python scripts/gemv_case.py 8192 28672 --copies 8 --chain | sed "s/^/WAVES=$wv /"
timeout 200 python scripts/gemv_case.py 8192 28672 --copies 8 | sed "s/^/WAVES=$wv /"
timeout 200 python scripts/gemv_case.py 4096 4096 --copies 64 | sed "s/^/WAVES=$wv /"
timeout 200 python scripts/gemv_case.py 8192 8192 --copies 16 --chain | sed "s/^/WAVES=$wv /"
done
