#!/bin/bash
# round 2, GPU session 3: whole suite with the MN-major-V attention (default), SKIP fix, L2-policy hints in VLAD
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r2_3_steps.log; }
: > gpurun_out/r2_3_steps.log
timeout 500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2_3_all.log 2>&1
stamp "all gpu tests (VMN=1): $(tail -1 gpurun_out/r2_3_all.log)"
ANYLOC_ATTN_VMN=0 timeout 200 python -m pytest tests/test_ops_gpu.py tests/test_vit_gpu.py -m gpu -q -k "attention or vs_oracle" > gpurun_out/r2_3_vmn0.log 2>&1
stamp "attention/vit tests (VMN=0): $(tail -1 gpurun_out/r2_3_vmn0.log)"
ANYLOC_ATTN_SKIP=1 timeout 300 python tools/diag_attn_skip.py > gpurun_out/r2_3_attn_skip.log 2>&1
stamp "attn skip diag: $(grep -c 'rel err' gpurun_out/r2_3_attn_skip.log) shapes"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_3_bench_vmn1.log 2>&1
stamp "bench VMN=1: $(grep -o '"value": [0-9.]*' gpurun_out/r2_3_bench_vmn1.log | head -1)"
ANYLOC_ATTN_VMN=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_3_bench_vmn0.log 2>&1
stamp "bench VMN=0: $(grep -o '"value": [0-9.]*' gpurun_out/r2_3_bench_vmn0.log | head -1)"
ANYLOC_ATTN_SKIP=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_3_bench_skip.log 2>&1
stamp "bench VMN=1 SKIP=1: $(grep -o '"value": [0-9.]*' gpurun_out/r2_3_bench_skip.log | head -1)"
timeout 100 python tools/diag_vlad.py > gpurun_out/r2_3_vlad_keep.log 2>&1
ANYLOC_VLAD_L2KEEP_MB=0 timeout 100 python tools/diag_vlad.py > gpurun_out/r2_3_vlad_nokeep.log 2>&1
stamp "vlad keep: $(grep -o 'L2-flushed [0-9.]* us' gpurun_out/r2_3_vlad_keep.log | tr '\n' ' ') | nokeep: $(grep -o 'L2-flushed [0-9.]* us' gpurun_out/r2_3_vlad_nokeep.log | tr '\n' ' ')"
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:vlad_ -c 12 --csv --log-file gpurun_out/r2_3_vlad_ncu_keep.csv python tools/diag_vlad.py --iters 1 > /dev/null 2>&1
ANYLOC_VLAD_L2KEEP_MB=0 timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:vlad_ -c 12 --csv --log-file gpurun_out/r2_3_vlad_ncu_nokeep.csv python tools/diag_vlad.py --iters 1 > /dev/null 2>&1
stamp "vlad ncu done"
cat gpurun_out/r2_3_steps.log
