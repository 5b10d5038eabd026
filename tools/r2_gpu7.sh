#!/bin/bash
# round 2, GPU session 7: re-scoring kernel v2, chunk-8 GEMM default, bench changes
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r2_7_steps.log; }
: > gpurun_out/r2_7_steps.log
timeout 400 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2_7_all.log 2>&1
stamp "all gpu tests: $(tail -1 gpurun_out/r2_7_all.log)"
timeout 200 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_7_bench_c3.log 2>&1
stamp "bench c3: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_7_bench_c3.log | head -1)"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2_7_bench_c2.log 2>&1
stamp "bench c2: $(grep -o '"value": [0-9.]*' gpurun_out/r2_7_bench_c2.log | head -1) $(grep -o '"features_rel_err": [0-9.e-]*' gpurun_out/r2_7_bench_c2.log)"
cat gpurun_out/r2_7_steps.log
