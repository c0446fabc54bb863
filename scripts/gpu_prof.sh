export ZG_BENCH_SHAPES=${SHAPES:-4096x14336} ZG_BENCH_FORMATS=${FMTS:-q8_0}
python bench.py --steps 2 --warmup 3 --no-cpu --rotation-mb 256 > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qgemv -s 10 -c 2 -f -o gpurun_out/${OUT:-prof} python bench.py --steps 2 --warmup 3 --no-cpu --rotation-mb 256 > gpurun_out/ncu_prof.log 2>&1
echo rc=$?
