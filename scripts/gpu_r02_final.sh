mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
HOT='regex:qgemv|attention|norm|chain|ew_mul|matmul|allreduce|k_elementwise|k_fused|k_rope|k_slice|k_head|k_repeat'
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$HOT" -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --gemv-steps 2 --no-cpu --no-extras > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/r02_launches_bench.csv
export EMULATE_WORLD=1 LAYERS=4 N_REPLAY=2
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k "$HOT" --launch-skip 120 -c 60 --csv --log-file gpurun_out/r02_launches_decode_70b_1gpu.csv python scripts/bench_sharded_emulate.py > gpurun_out/ncu_decode70.log 2>&1
wc -l gpurun_out/r02_launches_decode_70b_1gpu.csv
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('metric','value','unit','ms_per_step','e2e','gpu_launches','clocks')})
print(d['roofline']['kernel'][:60], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'])
print(d['gemv']['value'], [c['gbps'] for c in d['gemv']['cases']])
print(json.dumps(d.get('extras'))[:1800])
r=json.loads(open('gpurun_out/r02_bench_ref.json').read().strip().splitlines()[-1]); print({k:r.get(k) for k in ('impl','value','unit','ms_per_step')})
PY
