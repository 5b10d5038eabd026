#!/bin/bash
# round 2, last GPU session: sanitizer over the kernels changed since the first run, c1 / c5 / c3 bench lines, GEMM band A/B
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r2_final2_steps.log; }
: > gpurun_out/r2_final2_steps.log
bash tools/sanitize.sh > gpurun_out/r2_final2_sanitize.log 2>&1
stamp "sanitize: $(grep -c 'sanitize' gpurun_out/r2_final2_sanitize.log) runs"
cat gpurun_out/r2_final2_sanitize.log | tee -a gpurun_out/r2_final2_steps.log
timeout 200 python bench.py --workload c1 --steps 10 --warmup 3 --no-gpu-reference > gpurun_out/r2_final2_bench_c1.log 2>&1
stamp "bench c1: $(grep -o '"value": [0-9.]*' gpurun_out/r2_final2_bench_c1.log | head -1)"
timeout 200 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2_final2_bench_c5.log 2>&1
stamp "bench c5: $(grep -o '"value": [0-9.]*' gpurun_out/r2_final2_bench_c5.log | head -1)"
timeout 200 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_final2_bench_c3.log 2>&1
stamp "bench c3: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_final2_bench_c3.log | head -1)"
for b in 8 12 16; do
ANYLOC_GEMM_BAND=$b timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-parity-check > gpurun_out/r2_final2_bench_band$b.log 2>&1
stamp "band $b: $(grep -o '"value": [0-9.]*' gpurun_out/r2_final2_bench_band$b.log | head -1)"
done
cat gpurun_out/r2_final2_steps.log
