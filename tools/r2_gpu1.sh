#!/bin/bash
# round 2, GPU session 1: baseline suite + the untested tail-skipping attention variant (A/B in the bench)
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r2_1_steps.log; }
: > gpurun_out/r2_1_steps.log
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_1_all.log 2>&1
stamp "all gpu tests: $(tail -1 gpurun_out/r2_1_all.log)"
ANYLOC_ATTN_SKIP=1 timeout 120 python -m pytest tests/test_ops_gpu.py tests/test_vit_gpu.py -m gpu -x -q -k "attention or full_depth or vs_oracle" > gpurun_out/r2_1_skip.log 2>&1
stamp "attn skip tests: $(tail -1 gpurun_out/r2_1_skip.log)"
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_1_bench_base.log 2>&1
stamp "bench base: $(grep -o '"value": [0-9.]*' gpurun_out/r2_1_bench_base.log | head -1)"
ANYLOC_ATTN_SKIP=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_1_bench_skip.log 2>&1
stamp "bench skip: $(grep -o '"value": [0-9.]*' gpurun_out/r2_1_bench_skip.log | head -1)"
cat gpurun_out/r2_1_steps.log
