#!/bin/bash
# GPU box: SPACESHIP 640x360, 63 spp, loss off (pure tracing cost) under a few BVH / grid knobs
run() { echo "== $*"; env "$@" python tools/scene_bench.py --scenes scenes/spaceship-improved.npz --size 0 --budget 63 --extra bsdfSamplingFractionLoss=none 2>&1 | tail -1 ; }
run PPG_BVH_CT_X10=0
run PPG_BVH_CT_X10=10
run PPG_BVH_CT_X10=20
run PPG_BVH_CT_X10=30 PPG_BVH_LEAF=8
run PPG_BVH_CT_X10=10 PPG_BVH_LEAF=2
echo "== kl"; python tools/scene_bench.py --scenes scenes/spaceship-improved.npz --size 0 --budget 63 2>&1 | tail -1
