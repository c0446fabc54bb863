timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_a.log 2>&1; echo "bench rc=$?"
ZG_CUDA_BRANCH=15 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_br15.log 2>&1
ZG_CUDA_BRANCH=15 ZG_GEMV_NS=4 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_ns4.log 2>&1
ZG_CUDA_BRANCH=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_nobranch.log 2>&1
