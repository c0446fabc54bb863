# usage: bash scripts/gpu_r02_bench_n.sh N
set -x
N=$1
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "rc=$?"; tail -5 gpurun_out/r02_bench_n$N.err; cut -c1-2800 gpurun_out/r02_bench_n$N.json
