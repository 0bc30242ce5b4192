N=${1:-4}
TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
run() { echo "== $*"; env "$@" $TR bench.py --gpus $N --steps 15 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value']))"; }
for c in 16 64; do CG_PEER_CTAS=$c CG_DP_COMM=p2p $TR tools/dp_timeline.py 2>&1 | grep -E "^p2p" | tail -1; done
run CG_DP_COMM=p2p CG_PEER_CTAS=16
run CG_DP_COMM=p2p CG_PEER_CTAS=64
run CG_DP_COMM=nccl
