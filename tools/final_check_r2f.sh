#!/usr/bin/env bash
# last validation of the round: full GPU suite, reproducibility of the gradient-penalty path, one bench line
mkdir -p gpurun_out
timeout 60 python tools/flaky_check.py 2>&1 | tail -12 > gpurun_out/r2f_flaky_check.txt; grep ":" gpurun_out/r2f_flaky_check.txt | tail -6
timeout 300 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/r2f_gputests.log; tail -2 gpurun_out/r2f_gputests.log
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_paper.json 2>/dev/null; cut -c1-200 gpurun_out/r2f_bench_paper.json
