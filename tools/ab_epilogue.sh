#!/bin/bash
# A/B of the 2-CTA GEMM epilogue variants on one box (same clocks): staged vs direct stores, and the no-store ceiling.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$label.log 2>&1
  python - "$label" <<'PY' 2>&1 | tee -a gpurun_out/ab_summary.txt
import json, sys
label = sys.argv[1]
try:
    line = [l for l in open(f"gpurun_out/ab_{label}.log") if l.startswith("{")][-1]
    d = json.loads(line); r = d["roofline"]
    print(label, round(d["ms_per_step"], 2), "ms", round(d["value"], 1), "img/s shares", d["time_shares"],
          "gemm_ms", round(r["avg_launch_ms"] * r["launches"] / d["steps"], 2), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print(label, "FAILED", repr(e), open(f"gpurun_out/ab_{label}.log").read()[-400:])
PY
}
: > gpurun_out/ab_summary.txt
run staged     ANYLOC_GEMM_STAGED_EPI=1
run direct     ANYLOC_GEMM_STAGED_EPI=0
run staged2    ANYLOC_GEMM_STAGED_EPI=1
run st_skip_sw ANYLOC_GEMM_STAGED_EPI=1 ANYLOC_GEMM_DEBUG_SKIP_EPI=8
run st_skip_ls ANYLOC_GEMM_STAGED_EPI=1 ANYLOC_GEMM_DEBUG_SKIP_EPI=16
run st_skip_24 ANYLOC_GEMM_STAGED_EPI=1 ANYLOC_GEMM_DEBUG_SKIP_EPI=24
