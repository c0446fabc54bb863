set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -8 gpurun_out/r02_pytest_gpu.log
timeout 800 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "rc=$?"; tail -5 gpurun_out/r02_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print(d['value'], d['e2e']['value'], d['decode'])
print(json.dumps(d['extras'].get('prefill_llama3_8b_q8_0_2048')), json.dumps(d['extras'].get('decode_smollm_1p7b_q4_0_ctx512')))
for c in d['gemv']['cases']: print(c['K'],c['N'],c['format'],c['gbps'],c['frac_of_measured_hbm'])
PY
timeout 300 python scripts/trace_decode.py --model llama3-70b --kind q4_0 --context 512 --layers 6 --emulate-world 8 --show 30 > gpurun_out/r02_trace_70b_shard8.txt 2>&1; head -45 gpurun_out/r02_trace_70b_shard8.txt
