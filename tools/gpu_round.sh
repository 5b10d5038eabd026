#!/bin/bash
# One GPU session: VLAD parity + A/B timing + timeline + launch lists.
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/round_steps.log; }
: > gpurun_out/round_steps.log
timeout 400 python -m pytest tests/test_vlad_gpu.py -x -q > gpurun_out/t_vlad.log 2>&1; RC=$?
stamp "vlad tests rc=$RC: $(tail -1 gpurun_out/t_vlad.log)"
ANYLOC_VLAD=2 timeout 150 python tools/diag_vlad.py --save v2 --iters 5 > gpurun_out/diag_v2.log 2>&1
timeout 150 python tools/diag_vlad.py --compare v2 > gpurun_out/diag_v3.log 2>&1

ANYLOC_VLAD_TIMELINE=1 timeout 150 python tools/diag_vlad.py --iters 1 2>&1 | grep timeline | awk 'NR%5==1' > gpurun_out/timeline.log
stamp "diag: $(grep -c GB/s gpurun_out/diag_v3.log) v3 lines"
for sh in c2 c5; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ll_$sh.csv \
    python tools/diag_vlad.py --iters 2 --shape $sh > gpurun_out/ll_$sh.log 2>&1
done
stamp "launch lists done"
cat gpurun_out/round_steps.log
