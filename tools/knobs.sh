#!/bin/bash
# GPU box: runtime tuning knobs of libppg_b200.so (environment variables read by ppg_host.cu), one bench line each.
run() { echo "== $*"; env "$@" bash tools/ab.sh default; }
run PPG_NOP=1
run PPG_PIXEL_ORDER=1
run PPG_SMEM_CARVEOUT=25
run PPG_SMEM_CARVEOUT=50
run PPG_SMEM_CARVEOUT=100
run PPG_GRID_MULT=2
run PPG_GRID_MULT=4
run PPG_PIXEL_ORDER=1 PPG_GRID_MULT=2
