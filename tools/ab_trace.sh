#!/bin/bash
# GPU box: parity tests of the BVH scenes, then SPACESHIP with per-iteration host timing
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k 'spaceship or trace' 2>&1 | tail -n 12
PPG_TRACE=1 python bench.py --scene spaceship --steps 1 --warmup 1 --no-cpu-baseline --verbose > gpurun_out/bins_spaceship_v.json 2> gpurun_out/bins_spaceship_v.err
grep 'trace. iter' gpurun_out/bins_spaceship_v.err | cut -c1-230; grep -E "^\{'iteration|render_device" gpurun_out/bins_spaceship_v.err | cut -c1-260; cut -c1-130 gpurun_out/bins_spaceship_v.json
