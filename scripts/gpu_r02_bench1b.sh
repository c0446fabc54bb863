mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('metric','value','unit','ms_per_step','e2e','gpu_launches','clocks')})
print(d['roofline'])
print(d['gemv']['value'], [c['gbps'] for c in d['gemv']['cases']])
print(d.get('check')); print(d.get('extras'))
PY
