#!/bin/bash
# round 2, final GPU session: smoke(), whole suite, the bench lines committed under profiles/, GEMM chunk trade-off curve
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r2_final_steps.log; }
: > gpurun_out/r2_final_steps.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/r2_final_smoke.log 2>&1
stamp "smoke: $(grep smoke gpurun_out/r2_final_smoke.log | tail -1)"
timeout 400 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2_final_all.log 2>&1
stamp "all gpu tests: $(tail -1 gpurun_out/r2_final_all.log)"
timeout 300 python bench.py > gpurun_out/r2_final_bench_c2.log 2>&1
stamp "bench c2 (driver defaults): $(grep -o '"value": [0-9.]*' gpurun_out/r2_final_bench_c2.log | head -1)"
for c in 4 12 24; do
ANYLOC_GEMM_CHUNK=$c timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2_final_bench_chunk$c.log 2>&1
stamp "chunk $c: $(grep -o '"value": [0-9.]*' gpurun_out/r2_final_bench_chunk$c.log | head -1) $(grep -o '"features_rel_err": [0-9.e-]*' gpurun_out/r2_final_bench_chunk$c.log)"
done
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.log 2>&1
stamp "reference arm: $(grep -o '"value": [0-9.]*' gpurun_out/r2_final_bench_ref.log | head -1)"
cat gpurun_out/r2_final_steps.log
