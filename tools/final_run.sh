#!/usr/bin/env bash
# Round evidence on ONE B200 (run under gpurun): tests, the three bench configs, fp32 line, launch list, ncu --set full,
# role counters of the instrumented build. Everything lands in gpurun_out/r2_*.
T=${1:-r2}
python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/${T}_gputests.log; tail -2 gpurun_out/${T}_gputests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_paper.json 2> gpurun_out/${T}_bench_paper.err
python bench.py --config gp --steps 20 --warmup 5 > gpurun_out/${T}_bench_gp.json 2> gpurun_out/${T}_bench_gp.err
python bench.py --config scaled --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_scaled.json 2> gpurun_out/${T}_bench_scaled.err
python bench.py --fp32 --batch 16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_fp32_b16.json 2> gpurun_out/${T}_bench_fp32_b16.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
python tools/bench_layers.py --iters 30 > gpurun_out/${T}_layers.txt 2>&1
if [ -f tools/libinstr.so ]; then
  CG_LIB=tools/libinstr.so CG_TC_TIMING=1 python tools/bench_layers.py --iters 1 --only D1fwd,D1dgrad,D2fwd,D2dgrad,D3fwd,G5fwd 2>&1 | grep -E "tc3 timing|D conv|G conv" > gpurun_out/${T}_role_counters.txt
  CG_LIB=tools/libinstr.so CG_TC_TIMING=1 python tools/bench_layers.py --iters 1 --only D1wgrad,D2wgrad,D3wgrad,D4wgrad,D5wgrad 2>&1 | grep -E "wg2 timing" >> gpurun_out/${T}_role_counters.txt
  CG_LIB=tools/libinstr.so CG_TC_TIMING=1 python tools/bench_layers.py --batch 42 --iters 1 --only D1fwd,D2fwd 2>&1 | grep -E "tc3 timing|D conv" >> gpurun_out/${T}_role_counters.txt
fi
python tools/profile_step.py 128 2 > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py 128 2 > gpurun_out/${T}_ncu_time.log 2>&1
python tools/profile_step.py 128 1 > gpurun_out/${T}_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"rsgemm3_tc|wgrad2_tc|ghead_tc" -s 27 -c 27 -o gpurun_out/${T}_full python tools/profile_step.py 128 1 > gpurun_out/${T}_ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
ls -la gpurun_out/${T}_*
