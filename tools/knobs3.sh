#!/bin/bash
bash tools/ab.sh default build_variants/libppg_b512_m2.so build_variants/libppg_b1024_m1.so build_variants/libppg_c2.so
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "unguided or iteration_statistics or known_answers or full_render" 2>&1 | tail -2
