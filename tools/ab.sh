#!/bin/bash
# A/B helper for the GPU box: runs bench.py for each library variant given, prints value / ms / kernel ms.
for lib in "$@"; do
  if [ "$lib" = "default" ]; then unset PPG_B200_LIB; else export PPG_B200_LIB=$lib; fi
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']), round(d['ms_per_step'],1), round(d['kernel_ms_per_step'],1), 'bounce_avg_ms', round(d['roofline']['avg_launch_ms'],3), 'frac', round(d['roofline']['frac'],3))"
done
