#!/bin/bash
# round 2, GPU session 2: whole GPU suite with the new tests, new bench line (c2) and the c3 retrieval workload
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r2_2_steps.log; }
: > gpurun_out/r2_2_steps.log
timeout 420 python -m pytest tests -m gpu -q -x > gpurun_out/r2_2_all.log 2>&1
stamp "all gpu tests: $(tail -1 gpurun_out/r2_2_all.log)"
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_2_bench_c2.log 2>&1
stamp "bench c2: $(grep -o '"value": [0-9.]*' gpurun_out/r2_2_bench_c2.log | head -1)"
timeout 200 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/r2_2_bench_c3.log 2>&1
stamp "bench c3: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_2_bench_c3.log | head -1)"

ANYLOC_ATTN_SKIP=1 timeout 300 python tools/diag_attn_skip.py > gpurun_out/r2_2_attn_skip.log 2>&1
stamp "attn skip diag done"
cat gpurun_out/r2_2_steps.log
