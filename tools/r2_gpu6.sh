#!/bin/bash
# round 2, GPU session 6: many-tiles VLAD fix, coarse retrieval, GEMM chunk A/B, c3 / c5 bench lines
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r2_6_steps.log; }
: > gpurun_out/r2_6_steps.log
timeout 90 python -m pytest tests/test_vlad_gpu.py -m gpu -q -k "many_tiles" > gpurun_out/r2_6_many.log 2>&1
stamp "vlad many tiles: $(tail -1 gpurun_out/r2_6_many.log)"
timeout 120 python -m pytest tests/test_topk_gpu.py -m gpu -q > gpurun_out/r2_6_topk.log 2>&1
stamp "topk tests: $(tail -1 gpurun_out/r2_6_topk.log)"
timeout 400 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_vlad_gpu.py::test_vlad_many_tiles_per_cta > gpurun_out/r2_6_all.log 2>&1
stamp "all other gpu tests: $(tail -1 gpurun_out/r2_6_all.log)"
timeout 100 python tools/diag_vlad.py > gpurun_out/r2_6_vlad.log 2>&1
stamp "vlad: $(grep -o 'back-to-back [0-9.]* us' gpurun_out/r2_6_vlad.log | tr '\n' ' ') $(grep -o 'L2-flushed [0-9.]* us' gpurun_out/r2_6_vlad.log | tr '\n' ' ')"
timeout 200 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_6_bench_c3.log 2>&1
stamp "bench c3 coarse: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_6_bench_c3.log | head -1)"
ANYLOC_TOPK_COARSE=0 timeout 200 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_6_bench_c3_exact.log 2>&1
stamp "bench c3 exact: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_6_bench_c3_exact.log | head -1)"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_6_bench_c2.log 2>&1
stamp "bench c2: $(grep -o '"value": [0-9.]*' gpurun_out/r2_6_bench_c2.log | head -1)"
ANYLOC_GEMM_CHUNK=8 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_6_bench_c2_chunk8.log 2>&1
stamp "bench c2 chunk8: $(grep -o '"value": [0-9.]*' gpurun_out/r2_6_bench_c2_chunk8.log | head -1) $(grep -o '"features_rel_err": [0-9.e-]*' gpurun_out/r2_6_bench_c2_chunk8.log)"
timeout 200 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2_6_bench_c5.log 2>&1
stamp "bench c5: $(grep -o '"value": [0-9.]*' gpurun_out/r2_6_bench_c5.log | head -1)"
cat gpurun_out/r2_6_steps.log
