#!/usr/bin/env bash
# CTA-pair weight-gradient kernel: parity (layer + gradient tests), role counters, isolated launches and step time A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_layers_gpu.py tests/test_gradient_parity_gpu.py -m gpu -x -q > gpurun_out/wgp_tests.log 2>&1; echo "tests rc $?"
tail -5 gpurun_out/wgp_tests.log
CG_LIB=tools/libinstr.so CG_TC_TIMING=1 timeout 120 python tools/bench_layers.py --iters 1 --only D1wgrad,D2wgrad,D3wgrad,D4wgrad,D5wgrad 2>&1 | grep "wg2 timing" | tee gpurun_out/wgp_roles.txt | cut -c1-260
timeout 200 python tools/bench_layers.py 2>&1 | grep wgrad | tee gpurun_out/wgp_layers.txt
CG_WG_NO_PAIR=1 timeout 200 python tools/bench_layers.py 2>&1 | grep wgrad | tee gpurun_out/wgp_layers_nopair.txt
for i in 1 2; do
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/wgp_bench_$i.json | cut -c1-230
CG_WG_NO_PAIR=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/wgp_bench_nopair_$i.json | cut -c1-230
done
