#!/usr/bin/env bash
# 2-GPU check of the final build: DP equivalence (peer-memory exchange fp32 + bf16, NCCL bf16) and one bench line
TR="timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for cm in "p2p fp32" "p2p bf16" "nccl bf16"; do
  set -- $cm
  CG_DP_COMM=$1 $TR --master-port 29511 tools/dp_check.py $2 > gpurun_out/dpf_check_$1_$2.log 2>&1
  echo "[$1 $2] $(grep -E "rank 0|DP CHECK|Error|unavailable" gpurun_out/dpf_check_$1_$2.log | tail -2)"
done
$TR --master-port 29512 bench.py --gpus 2 --steps 15 --warmup 4 --no-cpu-baseline > gpurun_out/dpf_bench_n2.json 2> gpurun_out/dpf_bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/dpf_bench_n2.json').read().strip().splitlines()[-1]); print('N=2 ms/step %.3f value %.0f e2e %.0f' % (d['ms_per_step'], d['value'], d['e2e']['value']))"
