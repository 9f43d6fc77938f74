#!/bin/bash
run() { echo "== $*"; env "$@" bash tools/ab.sh default; }
run PPG_GRID_MULT=4
run PPG_GRID_MULT=8
run PPG_GRID_MULT=16
echo "== variants"; bash tools/ab.sh build_variants/libppg_b128_m8.so build_variants/libppg_b512_m2.so
