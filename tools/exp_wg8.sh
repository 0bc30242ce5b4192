#!/usr/bin/env bash
# bulk-reduce epilogue of the weight-gradient kernel: parity of the layer tests, role counters, layer times, step time A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_layers_gpu.py tests/test_gradient_parity_gpu.py -m gpu -x -q > gpurun_out/wg8_tests.log 2>&1; echo "tests rc $?" 
tail -3 gpurun_out/wg8_tests.log
CG_LIB=tools/libinstr.so CG_TC_TIMING=1 timeout 120 python tools/bench_layers.py --iters 1 --only D2wgrad,D3wgrad,D4wgrad,D5wgrad 2>&1 | grep "wg2 timing" | tee gpurun_out/wg8_roles.txt
timeout 200 python tools/bench_layers.py --only D1wgrad,D2wgrad,D3wgrad,D4wgrad,D5wgrad 2>&1 | tail -6 | tee gpurun_out/wg8_layers.txt
CG_WG_NO_BULK=1 timeout 200 python tools/bench_layers.py --only D1wgrad,D2wgrad,D3wgrad,D4wgrad,D5wgrad 2>&1 | tail -6 | tee gpurun_out/wg8_layers_nobulk.txt
for i in 1 2; do
timeout 200 python bench.py --steps 20 --warmup 5 2>/dev/null | tee gpurun_out/wg8_bench_$i.json | cut -c1-260
CG_WG_NO_BULK=1 timeout 200 python bench.py --steps 20 --warmup 5 2>/dev/null | tee gpurun_out/wg8_bench_nobulk_$i.json | cut -c1-260
done
