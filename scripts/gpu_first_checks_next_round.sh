# First GPU call of the next round: everything written after this round's GPU budget ended, each step under its own
# timeout so that a hang in the never-run CTA-pair kernel cannot take the box down (a hang there = kill + move on).
#   gpurun --timeout 420 -- 'bash scripts/gpu_first_checks_next_round.sh'
set -x
mkdir -p gpurun_out
# 1. the regular suite on the default paths (must stay green before anything experimental runs)
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
# 2. GGUF file -> HBM resident load (written after the last GPU run; only composes validated calls)
timeout 60 python -m pytest tests/test_zz_gpu_gguf_file.py -m gpu -q 2>&1 | tail -3
# 3. the CTA-pair prefill kernel (qgemm_cta2.cu), never executed so far: parity first, then throughput
ZG_GEMM_CTA2=1 timeout 60 python -m pytest tests/test_gpu_qmatmul.py -m gpu -x -q -k prefill > gpurun_out/cta2_tests.log 2>&1; echo "cta2 tests rc=$?"; tail -5 gpurun_out/cta2_tests.log
ZG_GEMM_CTA2=1 timeout 60 python scripts/bench_prefill.py --kind q8_0 --check --iters 12 > gpurun_out/cta2_prefill.log 2>&1; echo "cta2 bench rc=$?"; cat gpurun_out/cta2_prefill.log
timeout 60 python scripts/bench_prefill.py --kind q8_0 --check --iters 12 > gpurun_out/cta1_prefill.log 2>&1; cat gpurun_out/cta1_prefill.log
# 4. W8A8 gemv throughput (reference point of this round: 2.25 / 3.62 TB/s), then the never-run single-kernel form
timeout 60 python scripts/bench_w8a8.py > gpurun_out/w8a8.log 2>&1; cat gpurun_out/w8a8.log
ZG_W8A8_FUSED=1 timeout 60 python -m pytest tests/test_gpu_w8a8.py -m gpu -x -q 2>&1 | tail -3
ZG_W8A8_FUSED=1 timeout 60 python scripts/bench_w8a8.py > gpurun_out/w8a8_fused.log 2>&1; cat gpurun_out/w8a8_fused.log
nvidia-smi --query-gpu=clocks.sm,clocks_throttle_reasons.active --format=csv,noheader
