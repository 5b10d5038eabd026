#!/bin/bash
# One GPU session: VLAD parity + A/B timing + launch list + ncu full captures + bench + the whole GPU suite.
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/round_steps.log; }
: > gpurun_out/round_steps.log
timeout 400 python -m pytest tests/test_vlad_gpu.py -x -q > gpurun_out/t_vlad.log 2>&1; RC=$?
stamp "vlad tests rc=$RC: $(tail -1 gpurun_out/t_vlad.log)"
ANYLOC_VLAD=2 timeout 150 python tools/diag_vlad.py --save v2 > gpurun_out/diag_v2.log 2>&1
: > gpurun_out/diag_v3.log
for b in 1 2 4; do
  ANYLOC_VLAD_TMA_BURST=$b timeout 150 python tools/diag_vlad.py --compare v2 >> gpurun_out/diag_v3.log 2>&1
done
stamp "diag: $(grep -c GB/s gpurun_out/diag_v3.log) v3 lines"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vlad_launches.csv \
  python tools/diag_vlad.py --iters 2 > gpurun_out/ncu_diag.log 2>&1
stamp "ncu launch list done"
for sh in c2 c5; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:vlad_ -c 4 -f -o gpurun_out/prof_vlad3_$sh \
    python tools/diag_vlad.py --shape $sh --iters 1 > gpurun_out/ncu_full_$sh.log 2>&1
  ncu -i gpurun_out/prof_vlad3_$sh.ncu-rep --page raw --csv > gpurun_out/prof_vlad3_$sh.csv 2>/dev/null
done
stamp "ncu full captures done"
if [ $RC -eq 0 ]; then
  timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.log 2>&1
  stamp "bench: $(tail -c 400 gpurun_out/bench_c2.log | head -c 200)"
  timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_vlad_gpu.py > gpurun_out/t_all.log 2>&1
  stamp "all gpu tests: $(tail -1 gpurun_out/t_all.log)"
  timeout 200 python tools/diag_retrieval.py > gpurun_out/diag_retrieval.log 2>&1
  ANYLOC_TOPK_F16=0 timeout 200 python tools/diag_retrieval.py >> gpurun_out/diag_retrieval.log 2>&1
  stamp "retrieval: $(grep -c ms gpurun_out/diag_retrieval.log)"
fi
cat gpurun_out/round_steps.log
