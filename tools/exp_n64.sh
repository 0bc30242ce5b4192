export CG_LIB=tools/libinstr.so CG_TC_TIMING=1
for v in "" "CG_TC_DBG=1" "CG_TC_DBG=2" "CG_TC_DBG=4" "CG_TC_DBG=6" "CG_TC_DBG=7" "CG_TC_TPS=6" "CG_TC_TPS=4" "CG_TC_TPS=3"; do
  echo "== $v"; env $v python tools/bench_layers.py --iters 3 --only D1fwd,D2dgrad,D1dgrad 2>&1 | grep -E "tc3 timing|D conv" | awk 'NR%4==3 || NR%4==0'
done
