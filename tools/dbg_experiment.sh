for d in 0 1 2 4 6 7; do echo "== CG_TC_V1=1 CG_TC_DBG=$d"; CG_TC_V1=1 CG_TC_DBG=$d timeout 200 python tools/bench_layers.py --iters 10 2>&1 | grep -E "fwd|dgrad" | grep "D conv" ; done
