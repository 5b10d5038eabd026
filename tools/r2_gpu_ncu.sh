#!/bin/bash
# round 2: ncu captures for profiles/ -- launch list of the bench step, --set full of every GEMM mode + attention + LayerNorm,
# the VLAD kernels (c2, c5) and the retrieval kernels (c3)
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r2_ncu_steps.log; }
: > gpurun_out/r2_ncu_steps.log
NCU="ncu --clock-control none"
timeout 240 $NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file gpurun_out/r02_launches_c2.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference --no-parity-check --vocab random > gpurun_out/r2_ncu_launch.log 2>&1
stamp "launch list: $(wc -l < gpurun_out/r02_launches_c2.csv) lines"
timeout 240 $NCU --set full -k regex:'gemm_tc3|attention_tc16|layernorm' -c 16 -f -o gpurun_out/r02_ncu_vit python tools/diag_vit_once.py > gpurun_out/r2_ncu_vit.log 2>&1
stamp "vit full: $(tail -1 gpurun_out/r2_ncu_vit.log | cut -c1-80)"
timeout 200 $NCU --set full --import-source on -k regex:'vlad_' -s 6 -c 2 -f -o gpurun_out/r02_ncu_vlad_c2 python tools/diag_vlad.py --iters 1 --shape c2 > gpurun_out/r2_ncu_vlad_c2.log 2>&1
timeout 200 $NCU --set full -k regex:'vlad_' -s 6 -c 2 -f -o gpurun_out/r02_ncu_vlad_c5 python tools/diag_vlad.py --iters 1 --shape c5 > gpurun_out/r2_ncu_vlad_c5.log 2>&1
stamp "vlad full done"
NDB=10000 timeout 200 $NCU --set full -k regex:'gemm_tc3|topk_|normalize_rows' -s 14 -c 7 -f -o gpurun_out/r02_ncu_topk python tools/diag_retrieval.py > gpurun_out/r2_ncu_topk.log 2>&1
stamp "topk full: $(grep -c . gpurun_out/r2_ncu_topk.log) lines"
ls -la gpurun_out/*.ncu-rep | tee -a gpurun_out/r2_ncu_steps.log
for r in vit vlad_c2 vlad_c5 topk; do
  ncu -i gpurun_out/r02_ncu_$r.ncu-rep --page raw --csv > gpurun_out/r02_ncu_${r}_raw.csv 2>/dev/null
done
rm -f gpurun_out/r02_ncu_vit.ncu-rep gpurun_out/r02_ncu_vlad_c5.ncu-rep gpurun_out/r02_ncu_topk.ncu-rep   # keep the merge under 64 MiB
cat gpurun_out/r2_ncu_steps.log
