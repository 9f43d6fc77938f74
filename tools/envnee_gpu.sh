#!/bin/bash
# GPU box: environment light sampling against the oracle (diagnostics, then the tests), and the smoke render
mkdir -p gpurun_out
timeout 100 python tools/envnee_check.py > gpurun_out/r02_envnee_check.log 2>&1; echo "check rc=$?"; cut -c1-400 gpurun_out/r02_envnee_check.log | tail -n 45
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 100 python -m pytest tests/test_gpu_z_envnee.py -q -x 2>&1 | tail -n 25 > gpurun_out/r02_envnee_pytest.log; tail -n 12 gpurun_out/r02_envnee_pytest.log
