for v in "" "CG_TC_TPS=2" "CG_TC_TPS=3" "CG_TC_TPS=4" "CG_TC_TPS=6" "CG_TC_TPS=8" "CG_TC_TPS=12"; do
  echo "== $v"; env $v python tools/bench_layers.py --iters 30 2>&1 | grep -E "D conv|G convT" | grep -v wgrad | awk '{printf "%s%s%s %s | ", substr($1,1,1),$3,$4,$7} END{print ""}'
done
