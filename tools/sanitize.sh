#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over small-shape slices of the GPU suite: the kernels with
# hand-rolled mbarrier / cluster / ticket protocols are exactly where these tools earn their keep.
# Usage (GPU box): bash tools/sanitize.sh ; logs -> gpurun_out/sanitize_*.log (copy the summary into profiles/).
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL_OPS='(gemm_epilogues and tc3 and 128-256-64) or (gemm_epilogues and tc3 and 530-1152-384 and bias_split) or (test_attention and tc3 and (3-64-2 or 2-129-2)) or layernorm_split'
SEL_VLAD='golden or v3_odd_shapes or prepared_equals_plain or (many_tiles and 200-529-64-16)'
run() {  # tool, tag, pytest args...
  local tool=$1 tag=$2; shift 2
  timeout 600 $SAN --tool $tool --error-exitcode 99 --launch-timeout 0 python -m pytest "$@" -q -x -m gpu \
      > gpurun_out/sanitize_${tool}_${tag}.log 2>&1
  echo "[sanitize] $tool $tag: exit $? | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_${tag}.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitize_${tool}_${tag}.log | tail -1)"
}
for tool in memcheck racecheck; do
  run $tool ops tests/test_ops_gpu.py -k "$SEL_OPS"
  run $tool vlad tests/test_vlad_gpu.py -k "$SEL_VLAD"
  run $tool topk tests/test_topk_gpu.py -k "golden or coarse_pass_and_fallback or (vs_oracle and 2000-100-1024-20)"
done
