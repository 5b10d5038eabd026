#!/bin/bash
# Final GPU session of the round: whole GPU suite, bench line, launch list of the bench step, ncu full capture of the VLAD kernels.
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/round_steps.log; }
: > gpurun_out/round_steps.log
timeout 110 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1
stamp "all gpu tests: $(tail -1 gpurun_out/t_all.log)"
timeout 120 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.log 2>&1
stamp "bench: $(tail -c 300 gpurun_out/bench_c2.log | head -c 150)"
timeout 45 ncu --set full --clock-control none --import-source on -k regex:vlad_ -c 3 -f -o gpurun_out/prof_vlad3_final_c2 \
  python tools/diag_vlad.py --shape c2 --iters 1 > gpurun_out/ncu_full_c2.log 2>&1
ncu -i gpurun_out/prof_vlad3_final_c2.ncu-rep --page raw --csv > gpurun_out/prof_vlad3_final_c2.csv 2>/dev/null
stamp "ncu full c2 done"
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_final.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --vocab random > gpurun_out/ncu_bench.log 2>&1
stamp "bench launch list done"
cat gpurun_out/round_steps.log
