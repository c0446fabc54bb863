for v in 1 0 1 0; do
ZG_CUDA_AR_NORM=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$v bench.py --gpus 8 --no-extras --no-cpu --no-check --gemv-steps 1 --steps 30 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('AR_NORM=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
