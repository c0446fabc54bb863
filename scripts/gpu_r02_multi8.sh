# 8 GPUs: the 70B-width sharded parity at world 4 / 8 (peer, fused and NCCL paths), then the contract bench at N = 8 and 4.
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "peer-8 or fused-8 or nccl-8 or peer-4" > gpurun_out/r02_pytest_sharded_8gpu.log 2>&1; echo "sharded rc=$?"; tail -6 gpurun_out/r02_pytest_sharded_8gpu.log
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print('N=$N', d['value'], d['e2e']['value'], d['decode'], d['check'])
PY
done
