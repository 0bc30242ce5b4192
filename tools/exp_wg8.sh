#!/usr/bin/env bash
# bulk-reduce epilogue of the swapped weight-gradient mode: parity of the layer tests, role counters, layer times, step time A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_layers_gpu.py tests/test_gradient_parity_gpu.py -m gpu -x -q > gpurun_out/wg9_tests.log 2>&1; echo "tests rc $?" 
tail -3 gpurun_out/wg9_tests.log
CG_LIB=tools/libinstr.so CG_TC_TIMING=1 timeout 120 python tools/bench_layers.py --iters 1 --only D1wgrad 2>&1 | grep "wg2 timing" | tee gpurun_out/wg9_roles.txt
CG_WG_NO_BULK_SWAP=1 CG_LIB=tools/libinstr.so CG_TC_TIMING=1 timeout 120 python tools/bench_layers.py --iters 1 --only D1wgrad 2>&1 | grep "wg2 timing" | tee -a gpurun_out/wg9_roles.txt
timeout 200 python tools/bench_layers.py 2>&1 | grep wgrad | tee gpurun_out/wg9_layers.txt
CG_WG_NO_BULK_SWAP=1 timeout 200 python tools/bench_layers.py 2>&1 | grep wgrad | tee gpurun_out/wg9_layers_noswapbulk.txt
for i in 1 2; do
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/wg9_bench_$i.json | cut -c1-260
CG_WG_NO_BULK_SWAP=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/wg9_bench_noswapbulk_$i.json | cut -c1-260
done
