mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02_pytest_sharded_2gpu_b.log 2>&1; echo "sharded rc=$?"; tail -4 gpurun_out/r02_pytest_sharded_2gpu_b.log
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('metric','value','unit','ms_per_step','e2e','gpu_launches','scaling','n_gpus')})
print(d.get('check'))
PY
