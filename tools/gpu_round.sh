#!/bin/bash
# One GPU session: VLAD parity + A/B timing + assign-kernel limiter experiments + launch list + bench.
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/round_steps.log; }
: > gpurun_out/round_steps.log
timeout 400 python -m pytest tests/test_vlad_gpu.py -x -q > gpurun_out/t_vlad.log 2>&1; RC=$?
stamp "vlad tests rc=$RC: $(tail -1 gpurun_out/t_vlad.log)"
ANYLOC_VLAD=2 timeout 150 python tools/diag_vlad.py --save v2 --iters 5 > gpurun_out/diag_v2.log 2>&1
timeout 150 python tools/diag_vlad.py --compare v2 > gpurun_out/diag_v3.log 2>&1
stamp "diag: $(grep -c GB/s gpurun_out/diag_v3.log) v3 lines"
run_ll() {  # launch list of one configuration: $1 = tag, rest = env
  local tag=$1; shift
  env "$@" timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ll_$tag.csv \
    python tools/diag_vlad.py --iters 2 > gpurun_out/ll_$tag.log 2>&1
}
run_ll base X=1
run_ll st4 ANYLOC_VLAD_STAGES=4
run_ll st6 ANYLOC_VLAD_STAGES=6
run_ll nonorm ANYLOC_VLAD_DIAG=1
run_ll nomma ANYLOC_VLAD_DIAG=2
run_ll neither ANYLOC_VLAD_DIAG=3
stamp "launch lists done"
if [ $RC -eq 0 ]; then
  timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1
  stamp "bench: $(tail -c 400 gpurun_out/bench_c2.log | head -c 200)"
fi
cat gpurun_out/round_steps.log
