for v in "" "CG_TC_SST=2" "CG_TC_SST=2 CG_TC_TPS=5" "CG_TC_SST=2 CG_TC_TPS=6" "CG_TC_SST=4"; do
  echo "== $v"; env $v python tools/bench_layers.py --iters 30 --only D1fwd,D2dgrad,D2fwd,D3fwd 2>&1 | grep -E "D conv" | awk '{printf "%s%s%s %s | ", substr($1,1,1),$3,$4,$7} END{print ""}'
done
