#!/usr/bin/env bash
# Evidence for the last build of the round (CTA-pair weight-gradient kernel) on ONE B200, most important first.
T=r2e
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/${T}_gputests.log; tail -2 gpurun_out/${T}_gputests.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_paper.json 2> gpurun_out/${T}_bench_paper.err; cut -c1-200 gpurun_out/${T}_bench_paper.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 120 python tools/bench_layers.py --iters 30 > gpurun_out/${T}_layers.txt 2>&1
CG_LIB=tools/libinstr.so CG_TC_TIMING=1 timeout 120 python tools/bench_layers.py --iters 1 --only D1wgrad,D2wgrad,D3wgrad,D4wgrad,D5wgrad 2>&1 | grep -E "wg2 timing" > gpurun_out/${T}_role_counters_wgrad.txt
timeout 120 python tools/profile_step.py 128 2 > gpurun_out/${T}_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py 128 2 > gpurun_out/${T}_ncu_time.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"wgrad2p?_tc" -c 6 -o gpurun_out/${T}_wgrad python tools/profile_step.py 128 1 > gpurun_out/${T}_ncu_wgrad.log 2>&1
timeout 150 python bench.py --config scaled --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_scaled.json 2> gpurun_out/${T}_bench_scaled.err
timeout 150 python bench.py --config gp --steps 20 --warmup 5 > gpurun_out/${T}_bench_gp.json 2> gpurun_out/${T}_bench_gp.err
ls -la gpurun_out/${T}_*
