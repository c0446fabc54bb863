mkdir -p gpurun_out
export EMULATE_WORLD=1 LAYERS=2 N_REPLAY=2
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct \
  --clock-control none -k regex:"qgemv|attention|norm|chain|ewmul|allreduce" --launch-skip 72 -c 36 --csv --log-file gpurun_out/r02_ncu70_metrics.csv python scripts/bench_sharded_emulate.py > gpurun_out/ncu70.log 2>&1
tail -3 gpurun_out/ncu70.log; wc -l gpurun_out/r02_ncu70_metrics.csv
