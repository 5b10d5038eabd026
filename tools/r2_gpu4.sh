#!/bin/bash
# round 2, GPU session 4: suite with the re-scoring merged into the VLAD assignment kernel, SKIP=1 over the op/ViT tests,
# VLAD timing + true L2 residency (ncu --cache-control none), compute-sanitizer
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r2_4_steps.log; }
: > gpurun_out/r2_4_steps.log
timeout 500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2_4_all.log 2>&1
stamp "all gpu tests: $(tail -1 gpurun_out/r2_4_all.log)"
ANYLOC_ATTN_SKIP=1 timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_vit_gpu.py -m gpu -q --maxfail=5 > gpurun_out/r2_4_skip.log 2>&1
stamp "ops+vit tests (SKIP=1): $(tail -1 gpurun_out/r2_4_skip.log)"
timeout 100 python tools/diag_vlad.py > gpurun_out/r2_4_vlad.log 2>&1
stamp "vlad: $(grep -o 'back-to-back [0-9.]* us' gpurun_out/r2_4_vlad.log | tr '\n' ' ') $(grep -o 'L2-flushed [0-9.]* us' gpurun_out/r2_4_vlad.log | tr '\n' ' ')"
for keep in 76 0; do
ANYLOC_VLAD_L2KEEP_MB=$keep timeout 120 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:vlad_ -c 8 --csv --log-file gpurun_out/r2_4_vlad_ncu_nocc_keep$keep.csv python tools/diag_vlad.py --iters 1 --shape c2 > /dev/null 2>&1
done
stamp "vlad ncu (cache-control none) done"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_4_bench.log 2>&1
stamp "bench: $(grep -o '"value": [0-9.]*' gpurun_out/r2_4_bench.log | head -1)"
bash tools/sanitize.sh > gpurun_out/r2_4_sanitize.log 2>&1
stamp "sanitize: $(grep -c 'sanitize' gpurun_out/r2_4_sanitize.log) runs"
cat gpurun_out/r2_4_sanitize.log
cat gpurun_out/r2_4_steps.log
