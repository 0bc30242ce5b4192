#!/usr/bin/env bash
# usage (GPU box): bash tools/dp_run.sh N   -- DP equivalence checks + bench at N GPUs: peer-memory exchange vs NCCL
N=${1:-2}
TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for comm in p2p nccl; do
for m in fp32 bf16; do
  CG_DP_COMM=$comm $TR --master-port 29511 tools/dp_check.py $m > gpurun_out/dp_check_${comm}_$m.log 2>&1
  echo "[$comm $m] $(grep -E "rank 0|DP CHECK|Error|unavailable" gpurun_out/dp_check_${comm}_$m.log | tail -3)"
done
done
for comm in p2p nccl; do
  CG_DP_COMM=$comm $TR --master-port 29512 bench.py --gpus $N --steps 15 --warmup 4 --no-cpu-baseline > gpurun_out/dp_bench_n${N}_$comm.json 2> gpurun_out/dp_bench_n${N}_$comm.err
  grep -E "unavailable|Error" gpurun_out/dp_bench_n${N}_$comm.err | tail -2
  python - "gpurun_out/dp_bench_n${N}_$comm.json" $comm <<'P'
import sys, json
try:
  d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print('N=%d comm=%s ms/step %.3f value %.0f e2e %.0f streaming %.0f' % (d['n_gpus'], sys.argv[2], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e_streaming']['value']))
except Exception as e:
  print('no result:', e)
P
done
python bench.py --steps 15 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', round(d['ms_per_step'],3), round(d['value']))"
