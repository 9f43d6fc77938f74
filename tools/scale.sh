#!/bin/bash
# scaling sweep on one box (the driver runs the same commands): N = 1, 2, 4, 8
mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = "1" ]; then python bench.py --gpus 1 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/scale_$n.log 2>gpurun_out/scale_$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 3 --warmup 2 > gpurun_out/scale_$n.log 2>gpurun_out/scale_$n.err; fi
  tail -1 gpurun_out/scale_$n.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=$n', round(d['value']), 'Msamples/s', round(d['ms_per_step'],1), 'ms/step  e2e', round(d['e2e']['value']), 'var', d['final_variance'])" || tail -5 gpurun_out/scale_$n.err
done
