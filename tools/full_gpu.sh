#!/bin/bash
# GPU box: the whole GPU suite, then the bench lines of BASELINE configs 3-5
python -m pytest tests -m gpu -q 2>&1 | tail -n 25 > gpurun_out/r02_pytest_gpu.log; tail -n 8 gpurun_out/r02_pytest_gpu.log
for s in spaceship kitchen torus; do
  python bench.py --scene $s --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_$s.json 2> gpurun_out/r02_bench_$s.err
  cut -c1-160 gpurun_out/r02_bench_$s.json
done
