#!/bin/bash
# round 2, GPU session 5: suite with SKIP default + VLAD tail re-scoring / balanced tiles; VLAD timing; c2 and c3 bench lines
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r2_5_steps.log; }
: > gpurun_out/r2_5_steps.log
timeout 500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2_5_all.log 2>&1
stamp "all gpu tests: $(tail -1 gpurun_out/r2_5_all.log)"
timeout 100 python tools/diag_vlad.py > gpurun_out/r2_5_vlad.log 2>&1
stamp "vlad: $(grep -o 'back-to-back [0-9.]* us' gpurun_out/r2_5_vlad.log | tr '\n' ' ') $(grep -o 'L2-flushed [0-9.]* us' gpurun_out/r2_5_vlad.log | tr '\n' ' ')"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vlad_ -c 8 --csv --log-file gpurun_out/r2_5_vlad_ncu.csv python tools/diag_vlad.py --iters 1 > /dev/null 2>&1
stamp "vlad kernels: $(grep -o 'vlad_[a-z0-9_]*kernel[^,]*,[^,]*,[^,]*,[^,]*,[^,]*,[^,]*,[^,]*,"[0-9]*"' gpurun_out/r2_5_vlad_ncu.csv | sed 's/(.*,"/ /; s/"//' | tr '\n' ';' | cut -c1-400)"
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_5_bench_c2.log 2>&1
stamp "bench c2: $(grep -o '"value": [0-9.]*' gpurun_out/r2_5_bench_c2.log | head -1)"
timeout 200 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_5_bench_c3.log 2>&1
stamp "bench c3: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_5_bench_c3.log | head -1)"
timeout 200 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-parity-check > gpurun_out/r2_5_bench_c5.log 2>&1
stamp "bench c5: $(grep -o '"value": [0-9.]*' gpurun_out/r2_5_bench_c5.log | head -1)"
cat gpurun_out/r2_5_steps.log
