set -x
mkdir -p gpurun_out
timeout 800 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "rc=$?"; tail -5 gpurun_out/r02_bench_n1.err; python scripts/show_bench.py gpurun_out/r02_bench_n1.json 2>/dev/null | head -5; cut -c1-3000 gpurun_out/r02_bench_n1.json
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench_ref.err; cut -c1-1500 gpurun_out/r02_bench_ref.json
