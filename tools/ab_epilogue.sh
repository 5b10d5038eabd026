#!/bin/bash
# A/B of the 2-CTA GEMM epilogue variants on one box (same clocks): staged vs direct stores, and the no-store ceiling.
cd "$(dirname "$0")/.."
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('$label', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'img/s gemm_share', d['time_shares']['gemm_tc'], 'gemm_ms', round(r['avg_launch_ms']*r['launches']/d['steps'],2), 'clk', d['clocks']['sm_mhz'])"
}
run staged   ANYLOC_GEMM_STAGED_EPI=1
run direct   ANYLOC_GEMM_STAGED_EPI=0
run staged2  ANYLOC_GEMM_STAGED_EPI=1
run direct2  ANYLOC_GEMM_STAGED_EPI=0
run skip_qkv ANYLOC_GEMM_STAGED_EPI=0 ANYLOC_GEMM_DEBUG_SKIP_EPI=32
run skip_res ANYLOC_GEMM_STAGED_EPI=0 ANYLOC_GEMM_DEBUG_SKIP_EPI=16
run skip_all ANYLOC_GEMM_STAGED_EPI=0 ANYLOC_GEMM_DEBUG_SKIP_EPI=56
