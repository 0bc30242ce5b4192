N=${1:-2}
TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
run() { echo "== $*"; env "$@" $TR bench.py --gpus $N --steps 15 --warmup 4 --no-cpu-baseline $EXTRA 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value']))"; }
echo "== N=1"; python bench.py --steps 15 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value']))"
run CG_DP_COMM=none
run CG_DP_COMM=p2p
EXTRA=--no-dp-overlap run CG_DP_COMM=p2p
EXTRA= run CG_DP_COMM=nccl
EXTRA=--no-dp-overlap run CG_DP_COMM=none
