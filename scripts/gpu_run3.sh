timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
python scripts/bench_decode.py --model smollm-135m --kind q8_0 --tokens 64 --cpu-tokens 3 > gpurun_out/decode_135m.log 2>&1; tail -1 gpurun_out/decode_135m.log
python scripts/bench_decode.py --model smollm-1.7b --kind q4_0 --tokens 32 --context 512 > gpurun_out/decode_1p7b.log 2>&1; tail -1 gpurun_out/decode_1p7b.log
