N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
run() { echo "== $*"; env "$@" $TR bench.py --gpus $N --steps 15 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value']))"; }
echo "== N=1"; python bench.py --steps 15 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value']))"
echo "== N=1 CG_SM_LIMIT=140"; CG_SM_LIMIT=140 python bench.py --steps 15 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value']))"
run X=1
run NCCL_MAX_CTAS=2
run NCCL_MAX_CTAS=8
run CG_DP_BUCKETS=1
run CG_SM_LIMIT=140
run CG_SM_LIMIT=140 NCCL_MAX_CTAS=4
run CG_DP_BUCKETS=1 NCCL_MAX_CTAS=2
