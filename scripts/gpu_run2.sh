timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
python scripts/bench_prefill.py --kind q4_0 > gpurun_out/prefill_q4.log 2>&1; cat gpurun_out/prefill_q4.log | tail -3
