mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench_n8.err
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench_n4.err
python - <<'PY'
import json
for n in (8, 4):
    d=json.loads(open(f'gpurun_out/r02_bench_n{n}.json').read().strip().splitlines()[-1])
    print(n, {k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches','scaling','n_gpus')})
    print(d.get('check', {}).get('greedy_tokens'), d.get('check', {}).get('reduced_layers_vs_oracle'))
PY
