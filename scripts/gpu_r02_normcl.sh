timeout 600 python -m pytest tests/test_gpu_llama.py tests/test_gpu_synth_model.py tests/test_gpu_conformance.py -m gpu -x -q 2>&1 | tail -2
for v in 0 1; do
export ZG_CUDA_NORM_CLUSTER=$v
for wd in 1 8; do EMULATE_WORLD=$wd LAYERS=8 timeout 300 python scripts/bench_sharded_emulate.py 2>&1 | tail -1 | cut -c1-90 | sed "s/^/CLUSTER=$v /"; done
done
ZG_CUDA_NORM_CLUSTER=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --gemv-steps 1 --decode-layers 2 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('check', d['check']['reduced_layers_vs_oracle'])"
ZG_CUDA_NORM_CLUSTER=1 timeout 200 python scripts/trace_decode.py --model llama3-70b --kind q4_0 --context 512 --layers 4 --emulate-world 8 --show 4 2>&1 | grep -E "rmsnorm|kernels"
