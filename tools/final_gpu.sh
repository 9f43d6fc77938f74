#!/bin/bash
# GPU box, last call of round 2 (seconds of budget): the smoke render, the plain C host, and SPACESHIP against the oracle -- i.e. the host-side changes made
# after the last full GPU run (parallel scene set-up, exception guards) on a scene that uses them (457 560 triangles through the parallel BVH build / packing).
mkdir -p gpurun_out
{ timeout 30 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
  timeout 60 python -m pytest tests/test_gpu_z_envnee.py tests/test_gpu_parity.py -q -x -k "plain_c or spaceship_matches_oracle" 2>&1 | tail -n 6; } | tee gpurun_out/r02_final_check.log
