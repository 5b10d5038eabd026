"""CPU test of bench.py's reference arm (the oracle port timed on host cores) and its bookkeeping."""
import json
import os
import subprocess
import sys

from tests.util import ROOT


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "1", "--warmup", "1", "--ref-images", "1"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "cpu_baseline",
                "e2e"):
        assert key in line
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_flops_model_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    # c2 987.3 GFLOP/img, c5 847.8, c1 9.29 (early-exit form)
    assert abs(bench.vit_flops_per_image("dinov2_vitg14", 31, 322, 322) / 1e9 - 987.3) < 1.0
    assert abs(bench.vit_flops_per_image("dinov2_vitl14", 20, 518, 518) / 1e9 - 847.8) < 1.0
    assert abs(bench.vit_flops_per_image("dinov2_vits14", 9, 224, 224) / 1e9 - 9.29) < 0.05
    assert bench.usable_cores() >= 1


def test_roofline_traffic_comes_from_committed_ncu_exports():
    """bench.py's `roofline.traffic` is read from the ncu --set full exports under profiles/, not hard-coded."""
    sys.path.insert(0, ROOT)
    import bench
    gemm, src = bench.ncu_traffic(["kernel<1, 1, 1>", "kernel<1, 3, 1>", "kernel<1, 4, 1>"], "vit")
    assert src and src.startswith("profiles/") and 3e8 < gemm < 1e9          # mean of the four per-block GEMMs (MB range)
    vlad, src = bench.ncu_traffic(["vlad_assign_tc_kernel", "vlad_accumulate3_kernel"], "vlad_c2", per_call=True)
    assert src and 1.1e8 < vlad < 3e8                                         # >= the 110.5 MB algorithmic bytes of c2
    assert bench.ncu_traffic(["no_such_kernel"], "vit") == (None, None)
