#!/usr/bin/env bash
# usage (GPU box): bash tools/dp_run.sh N   -- DP equivalence checks + bench at N GPUs with / without the overlap
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for m in fp32 bf16; do
  $TR --master-port 29511 tools/dp_check.py $m > gpurun_out/dp_check_$m.log 2>&1
  grep -E "rank 0|DP CHECK|Error" gpurun_out/dp_check_$m.log | tail -3
done
for f in "" "--no-dp-overlap"; do
  $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 $f > gpurun_out/dp_bench_n${N}${f}.json 2> gpurun_out/dp_bench_n${N}${f}.err
  tail -c 400 gpurun_out/dp_bench_n${N}${f}.err | grep -v Warning | tail -3
  python - "gpurun_out/dp_bench_n${N}${f}.json" <<'P'
import sys, json
try:
  d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print('N=%d overlap=%s ms/step %.3f value %.0f e2e %.0f streaming %.0f' % (d['n_gpus'], d['config']['dp_overlap'], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e_streaming']['value']))
except Exception as e:
  print('no result:', e)
P
done
