mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02_pytest_sharded_arnorm.log 2>&1; echo "sharded rc=$?"; tail -4 gpurun_out/r02_pytest_sharded_arnorm.log
for v in 1 0; do
ZG_CUDA_AR_NORM=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$v bench.py --gpus 2 --no-extras --no-cpu --gemv-steps 2 --steps 30 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('AR_NORM=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['check']['greedy_tokens'][:3], d['check']['reduced_layers_vs_oracle']['logits_rel_err_vs_oracle'])"
done
