mkdir -p gpurun_out
for st in 0 2; do
export ZG_GEMV_STREAM=$st
echo "== STREAM=$st"
timeout 200 python scripts/gemv_case.py 4096 14336 --copies 32
timeout 200 python scripts/gemv_case.py 4096 4096 --copies 64
timeout 200 python scripts/gemv_case.py 8192 28672 --copies 8 --chain
timeout 200 python scripts/gemv_case.py 8192 8192 --copies 16 --chain
timeout 200 python scripts/gemv_case.py 8192 28672 --copies 8
done
for st in 0 2; do
ZG_GEMV_STREAM=$st timeout 300 ncu --set full --import-source on --clock-control none -k regex:qgemv --launch-skip 6 -c 1 -o gpurun_out/r02_q4_4096x14336_stream$st python scripts/gemv_case.py 4096 14336 --copies 32 --reps 2 > gpurun_out/ncu_case_$st.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
