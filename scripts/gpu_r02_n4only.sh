mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --no-extras > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "rc=$?"; tail -2 gpurun_out/r02_bench_n4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n4.json').read().strip().splitlines()[-1])
print(4, {k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches','scaling','n_gpus')})
print(d.get('check', {}).get('greedy_tokens'), d.get('check', {}).get('reduced_layers_vs_oracle'))
PY
