"""GPU diagnostic: full-depth extractor error vs the CPU oracle for both GEMM engines."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anyloc_b200 import utilities as u
from oracle import anyloc_oracle as ao, dinov2_restated as dr

for name, depth, layer, HW in (("dinov2_vitg14", 32, 31, 98), ("dinov2_vitl14", 21, 20, 98)):
    model = dr.build(name, seed=0, depth_override=depth)
    img = torch.randn(2, 3, HW, HW, generator=torch.Generator().manual_seed(1234))
    t = time.time(); ref = ao.extract_features(model, img, layer, "value"); tc = time.time() - t
    ref64 = ao.extract_features(model.double(), img.double(), layer, "value")
    print(f"{name} L{layer} {HW}x{HW}: cpu oracle {tc:.1f}s; fp32-oracle vs fp64-oracle err {float((ref.double()-ref64).abs().max()/ref64.abs().max()):.2e}")
    model = model.float()
    for eng in ("simt", "tc3"):
        ext = u.DinoV2ExtractFeatures(name, layer, "value", device="cuda", weights=model.state_dict(), gemm_engine=eng)
        out = ext(img.cuda()).cpu()
        print(f"   {eng}: vs fp32 oracle {float((out-ref).abs().max()/ref.abs().max()):.2e}  vs fp64 oracle {float((out.double()-ref64).abs().max()/ref64.abs().max()):.2e}")
