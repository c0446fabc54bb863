# ncu evidence for profiles/: (1) launch list of the default bench command, (2) DRAM traffic per GEMV case,
# (3) one --set full capture of the dominant GEMV kernel, (4) tensor-pipe metrics of the prefill GEMM.
set -x
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
for shp in 4096x4096 4096x14336; do for fmt in i8_f32 q8_0 q4_0; do
  ZG_BENCH_SHAPES=$shp ZG_BENCH_FORMATS=$fmt ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:qgemv -s 8 -c 6 --csv --log-file gpurun_out/r01_traffic_${shp}_${fmt}.csv python bench.py --steps 2 --warmup 3 --no-cpu --rotation-mb 256 > /dev/null 2>&1
done; done
ZG_BENCH_SHAPES=4096x4096 ZG_BENCH_FORMATS=q4_0 ncu --set full --clock-control none --import-source on -k regex:qgemv -s 8 -c 2 -f -o gpurun_out/r01_qgemv_q4_full python bench.py --steps 2 --warmup 3 --no-cpu --rotation-mb 256 > /dev/null 2>&1
python scripts/bench_prefill.py --shapes 4096x4096 --iters 4 > gpurun_out/plain_prefill.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:qgemm -s 2 -c 3 --csv --log-file gpurun_out/r01_qgemm_tensor.csv python scripts/bench_prefill.py --shapes 4096x4096 --iters 4 > /dev/null 2>&1
echo done
