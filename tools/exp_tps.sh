for v in "" "CG_TC_TPS=7" "CG_TC_TPS=3" "CG_TC_TPS=2" "CG_TC_TPS=6"; do
  echo "== $v"; env $v python tools/bench_layers.py --iters 30 --only D1fwd 2>&1 | grep -E "D conv"
done
