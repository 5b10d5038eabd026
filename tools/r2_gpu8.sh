#!/bin/bash
# round 2, GPU session 8: single-pass index build, bench traffic from ncu exports, one-step launch list
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r2_8_steps.log; }
: > gpurun_out/r2_8_steps.log
timeout 400 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2_8_all.log 2>&1
stamp "all gpu tests: $(tail -1 gpurun_out/r2_8_all.log)"
timeout 200 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_8_bench_c3.log 2>&1
stamp "bench c3: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_8_bench_c3.log | head -1) build $(grep -o '"index_build_local_shard_ms": [0-9.]*' gpurun_out/r2_8_bench_c3.log)"
ANYLOC_BENCH_PROFILE_STEP=1 timeout 240 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c2_step.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-parity-check > gpurun_out/r2_8_launch.log 2>&1
stamp "launch list: $(wc -l < gpurun_out/r02_launches_c2_step.csv) lines"
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_8_bench_c2.log 2>&1
stamp "bench c2: $(grep -o '"value": [0-9.]*' gpurun_out/r2_8_bench_c2.log | head -1)"
cat gpurun_out/r2_8_steps.log
