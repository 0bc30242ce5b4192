for d in 0 8 16 24; do echo "== DBG=$d"; CG_TC_V1=1 CG_TC_DBG=$d CG_TC_TIMING=1 timeout 200 python tools/bench_layers.py --iters 1 2>&1 | grep -E "tc timing" | awk 'NR%2==0' | head -3; done
