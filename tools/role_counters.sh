#!/usr/bin/env bash
# Per-role cycle counters of the CTA-pair kernel (CTA 0): waits of the MMA-issuing warp on accumulator-empty / slab /
# weight barriers, MMA issue and commit time, epilogue wait and work. Needs the instrumented build:
#   CG_TC_INSTRUMENT=1 bash calciumgan_b200/csrc/build.sh && gpurun -- 'bash tools/role_counters.sh'
# Timing experiments on the same kernel: CG_TC_DBG bits (1 no epilogue, 2 no weight loads, 4 no activation loads),
# CG_TC_NK=1..4 (K steps issued per 64-channel chunk; results are wrong, only the time is of interest).
CG_TC_TIMING=1 timeout 300 python tools/bench_layers.py --iters 1 2>&1 | grep -E "tc3 timing"
