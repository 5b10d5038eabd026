#!/usr/bin/env python
"""bench.py -- images/s of the DINOv2 -> hard-VLAD descriptor pipeline (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A step = one pass of the hot path over one batch of synthetic images per GPU:
DinoV2ExtractFeatures.__call__ (ViT forward, early exit at the hooked layer) -> VLAD.generate_multi.
Workload at N=1 is BASELINE.json configs[1]: ViT-G/14 layer-31 'value', 322x322, K=32, batch 32.
Under torchrun (N>1) every rank runs the same per-GPU batch on its own images (weak scaling, no
data-path collective inside the step; the descriptor all-gather belongs to the retrieval configs).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "c1": dict(model="dinov2_vits14", layer=9, facet="value", H=224, W=224, K=8, B=16,
               name="c1: ViT-S/14 layer-9 value, 224x224, K=8 VLAD, batch 16"),
    "c2": dict(model="dinov2_vitg14", layer=31, facet="value", H=322, W=322, K=32, B=32,
               name="c2: ViT-G/14 layer-31 value, 322x322, K=32 VLAD, batch 32"),
    "c5": dict(model="dinov2_vitl14", layer=20, facet="value", H=518, W=518, K=128, B=64,
               name="c5: ViT-L/14 layer-20 value, 518x518, K=128 VLAD, batch 64"),
}
METRIC = "images/sec end-to-end DINOv2-VLAD descriptors"
UNIT = "images/s"


def vit_flops_per_image(model, layer, H, W):
    """2*N*588*D + L*(24*T*D^2 + 4*T^2*D) + 2*T*D^2 (early-exit form)."""
    from anyloc_b200.vit import ARCHS, ffn_hidden
    D, _, _, kind = ARCHS[model]
    N = (H // 14) * (W // 14)
    T = N + 1
    hid = ffn_hidden(D, kind)
    ffn = (2 * T * D * 2 * hid + 2 * T * hid * D) if kind != "mlp" else 4 * T * D * hid
    per_block = 2 * T * D * 3 * D + 2 * T * D * D + 4 * T * T * D + ffn
    return 2 * N * 588 * D + layer * per_block + 2 * T * D * D


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tflops_burst=p["bf16_tflops"],
                    tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


def usable_cores():
    """Host threads this process can really use: min(affinity, cgroup CPU quota).  (On the GPU boxes
    os.cpu_count() says 128 while cpu.max grants 16 CPUs; 128 threads then run 40x slower.)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // p))
        except Exception:
            pass
    return n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ reference arm / cpu baseline
def cpu_reference(wl, n_images, steps, warmup, seed=0):
    """The reference's own CPU path, restated (oracle/): per image, batch 1, the FULL model forward
    with the facet hook (scripts/dino_v2_vlad.py:164-188 -> utilities.py:263-285), then
    VLAD.generate per image with the [N,K,D] residual tensor (utilities.py:819-890, :956-962).
    Returns (images_per_s, ms_per_step, cores)."""
    import numpy as np
    import torch
    from oracle import anyloc_oracle as ao
    from oracle import dinov2_restated as dr
    cores = usable_cores()
    torch.set_num_threads(cores)
    model = dr.build(wl["model"], seed=seed)                 # all blocks: the reference runs them all
    D = model.embed_dim
    g = torch.Generator().manual_seed(1234)
    imgs = torch.randn(n_images, 3, wl["H"], wl["W"], generator=g)
    centers = 0.6 * torch.nn.functional.normalize(torch.randn(wl["K"], D, generator=g), dim=1)

    def one_step():
        feats = [ao.extract_features_full_forward(model, imgs[i:i + 1], wl["layer"], wl["facet"]) for i in range(n_images)]
        feats = torch.cat(feats)
        return torch.stack([ao.vlad_generate_faithful(f, centers) for f in feats])

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    return n_images * steps / dt, dt / steps * 1e3, cores


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_images
    ips, ms, cores = cpu_reference(wl, n, args.steps, max(args.warmup, 1))
    sample = f"{n} of {wl['B']} images per step, batch 1 per image, all blocks + hook, CPU VLAD with [N,K,D] residuals"
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "sample": sample},
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ this repo
def run_ours(args, wl):
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (anyloc_b200 has no CPU fallback); use --impl reference")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from anyloc_b200 import _lib, utilities as u
    from anyloc_b200.vit import random_state_dict, ARCHS

    B, H, W, K = wl["B"], wl["H"], wl["W"], wl["K"]
    D = ARCHS[wl["model"]][0]
    sd = random_state_dict(wl["model"], seed=0, device=dev, depth=wl["layer"] + 1)
    ext = u.DinoV2ExtractFeatures(wl["model"], wl["layer"], wl["facet"], device=dev, weights=sd,
                                  gemm_engine=args.engine, precision=args.precision)
    del sd
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    img_dev = torch.randn(B, 3, H, W, device=dev, generator=g)
    img_host = img_dev.cpu().pin_memory()
    feats = ext(img_dev)
    np.random.seed(42)
    vlad = u.VLAD(K)
    if args.vocab == "fit":
        vlad.fit(feats.reshape(-1, D))      # vocabulary on this batch's features (GPU k-means)
    else:                                   # profiling runs: skip the k-means launches
        vlad.kmeans = u._KMeans(K, mode="cosine")
        idx = torch.randperm(feats.shape[0] * feats.shape[1], device=dev, generator=g)[:K]
        vlad.c_centers = vlad.kmeans.centroids = 0.7 * feats.reshape(-1, D)[idx].contiguous()
        vlad.desc_dim = D
    out_host = [torch.empty(B, K * D, dtype=torch.float32).pin_memory() for _ in range(2)]

    def step_device():
        return vlad.generate_multi(ext(img_dev))

    def step_e2e(i):
        x = img_host.to(dev, non_blocking=True)                    # H2D of this step's inputs
        out_host[i & 1].copy_(vlad.generate_multi(ext(x)), non_blocking=True)   # D2H of the result

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # end-to-end through the public API with host buffers (pinned), copies inside the timed region
    for i in range(max(1, args.warmup // 2)):
        step_e2e(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_e2e(i)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    g_ms, g_n, g_fl = prof["gemm_tc"]
    flops_img = vit_flops_per_image(wl["model"], wl["layer"], H, W)
    roof = None
    if g_n:
        ach = g_fl / (g_ms / 1e3) / 1e12
        f16 = args.precision == "f16x3"
        passes = 3.0 if f16 else 6.0       # bf16-rate-equivalent tensor passes per algorithmic product
        two_cta = os.environ.get("ANYLOC_GEMM_2CTA", "1") != "0"
        kname = ("gemm_tc3_2cta_kernel<%s> (tcgen05 cta_group::2 M256xN256, kind::%s" if two_cta else
                 "gemm_tc3_kernel<256,%s> (tcgen05 cta_group::1 M128xN256, kind::%s") % (
                     "true" if f16 else "false", "f16" if f16 else "tf32")
        # DRAM bytes per launch of this kernel from the committed ncu --set full capture of the same command
        # (profiles/r01_ncu_summary.md, "Final state": mean over the four per-block GEMMs w3/qkv/proj/w12); only
        # valid for the configuration that was captured
        traffic = 617.5e6 if (args.workload == "c2" and f16 and two_cta) else None
        roof = {"kernel": kname + ", 3-term split, fp32 accumulate, RN chunk accumulation)",
                "bound": "tensor", "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["tflops_sustained"], "traffic": traffic,
                "traffic_source": "profiles/r01_ncu_summary.md (ncu --set full, dram__bytes_read+write per launch)" if traffic else None,
                "algorithmic_bytes_per_launch": 425.5e6 if traffic else None,
                "peak_source": f"{peaks['source']} cuBLAS bf16 sustained (MEASURED_PEAKS.json)",
                "note": "achieved = algorithmic 2MNK FLOPs / device time; the engine issues 3 MMAs per product "
                        "(fp32-equivalent accuracy) at %s the bf16 rate, i.e. %d bf16-equivalent passes"
                        % ("1x" if f16 else "0.5x", int(passes)),
                "tensor_pipe_frac_est": passes * ach / peaks["tflops_sustained"],
                "launches": g_n, "avg_launch_ms": g_ms / g_n, "share_of_step": g_ms / ms_total}
    v_ms, v_n, v_bytes = prof["vlad"]
    vroof = None
    if v_n:
        gbs = v_bytes / (v_ms / 1e3) / 1e9
        vroof = {"kernel": "VLAD v3: vlad_assign_tc_kernel (TMA + tcgen05 coarse scores + row norms) -> "
                           "vlad_rescore_amb_kernel (ambiguous rows only) -> vlad_accumulate3_kernel (+ fused "
                           "normalisation); prepared vocabulary, 3 launches",
                 "bound": "hbm", "achieved": gbs,
                 "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                 # ncu --set full, c2 shape (profiles/r01_vlad_v3.md): DRAM read+write of the three launches
                 "traffic": 235.0e6 if args.workload == "c2" else None,
                 "algorithmic_bytes_per_launch_group": v_bytes / v_n,
                 "avg_launch_ms": v_ms / v_n, "share_of_step": v_ms / ms_total}
    shares = {c: round(prof[c][0] / ms_total, 4) for c in prof if prof[c][1]}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32-equivalent (tcgen05 %s 3-term split, fp32 accumulate)" % ("fp16" if args.precision == "f16x3" else "tf32"),
            "data": "synthetic",
            "config": {"workload": wl["name"], "per_gpu_batch": B, "global_batch": B * world,
                       "weights": "random-init (upstream recipe), no checkpoint offline",
                       "parallelism": f"dp{world} (images sharded, no collective in the step)",
                       "precision": args.precision,
                       "cache": "inputs larger than L2: weights (hi+lo) streamed every step"},
            "vit_tflops_algorithmic": flops_img * value / 1e12,
            "roofline": roof, "roofline_vlad": vroof, "time_shares": shares,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * img_host.numel() * 4,
                    "d2h_bytes_per_step": world * B * K * D * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks}
    if world == 1 and not args.no_cpu_baseline:
        n_ref = max(args.ref_images, 4)        # ~14 s of timed CPU work at c2 (0.85-0.9 img/s on the box's 16 usable cores)
        ips, ms, cores = cpu_reference(wl, n_ref, 3, 1)
        line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n_ref} images x 3 steps of the same workload, batch 1 per image, "
                                          "all blocks + hook, CPU VLAD with [N,K,D] residuals"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--engine", default="auto", choices=["auto", "tc3", "simt"])
    ap.add_argument("--ref-images", type=int, default=2, help="images per CPU-reference step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--vocab", default="fit", choices=["fit", "random"])
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "tf32x3"],
                    help="operand pair format of the tensor-core GEMMs (both fp32-equivalent; see DESIGN.md)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
