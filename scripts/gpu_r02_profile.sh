# ncu evidence for profiles/ (round 2): (1) launch list of the bench command, (2) DRAM traffic per GEMV case,
# (3) one --set full capture of the streamed matvec inside a Llama-3-70B decode step, (4) launch list of that step.
mkdir -p gpurun_out
HOT='regex:qgemv|attention|norm|chain|ewmul|matmul|allreduce|k_ew|k_rope|k_slice|k_batched'
python bench.py --steps 2 --warmup 3 --gemv-steps 2 --no-cpu --no-extras > gpurun_out/plain_bench.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain_bench.log; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$HOT" -c 5000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --gemv-steps 2 --no-cpu --no-extras > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/r02_launches_bench.csv
for shp in 4096x4096 4096x14336; do for fmt in i8_f32 q8_0 q4_0; do
  ZG_BENCH_SHAPES=$shp ZG_BENCH_FORMATS=$fmt timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:qgemv -s 8 -c 6 --csv --log-file gpurun_out/r02_traffic_${shp}_${fmt}.csv python bench.py --steps 2 --warmup 3 --gemv-steps 2 --no-cpu --no-extras --no-check --decode-layers 1 --rotation-mb 256 > /dev/null 2>&1
done; done
export EMULATE_WORLD=1 LAYERS=4 N_REPLAY=2
python scripts/bench_sharded_emulate.py > gpurun_out/plain_emulate.log 2>&1 || { echo "plain emulate failed"; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qgemv_stream -s 16 -c 3 -f -o gpurun_out/r02_qgemv_stream_q4_70b_full python scripts/bench_sharded_emulate.py > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k "$HOT" --launch-skip 108 -c 54 --csv --log-file gpurun_out/r02_launches_decode_70b_1gpu.csv python scripts/bench_sharded_emulate.py > gpurun_out/ncu_decode70.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_*.csv
