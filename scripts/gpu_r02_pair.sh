set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_llama.py tests/test_gpu_synth_model.py tests/test_gpu_conformance.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -5
for P in 0 1; do
  for M in smollm-135m smollm-1.7b; do
    ZG_CUDA_GEMV_PAIR=$P timeout 300 python scripts/bench_decode.py --model $M --kind q4_0 --context 512 --tokens 64 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PAIR=$P', d['model'], d['device_tok_s'], d['value'], d['kernels_per_token'])"
  done
  ZG_CUDA_GEMV_PAIR=$P timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -4
done
