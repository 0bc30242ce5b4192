#!/usr/bin/env bash
# usage (8-GPU box): bash tools/scale_run.sh [tag]  -- bench.py at 1/2/4/8 GPUs of the same box -> gpurun_out/<tag>_scale_n*.json
TAG=${1:-r2}
for N in 1 2 4 8; do
  if [ $N = 1 ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py"; fi
  $CMD --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
  python - gpurun_out/${TAG}_scale_n$N.json <<'P'
import sys, json
try:
  d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print('N=%d ms/step %.3f value %.0f e2e %.0f streaming %.0f clocks %s' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e_streaming']['value'], d['clocks']))
except Exception as e:
  print('no result:', e)
P
done
