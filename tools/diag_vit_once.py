"""One ViT-G extractor call at the c2 batch shape (B=32, 322x322) with 3 blocks loaded (layer 2, value facet): the per-block
kernels have exactly the shapes of the 31-block bench step, so an `ncu --set full` capture of this command profiles
qkv / attention / proj / w12 / w3 / LayerNorm as the step runs them, without replaying 31 blocks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anyloc_b200 import utilities as u
from anyloc_b200.vit import random_state_dict

sd = random_state_dict("dinov2_vitg14", seed=0, device="cuda", depth=3)
ext = u.DinoV2ExtractFeatures("dinov2_vitg14", 2, "value", device="cuda", weights=sd, precision=os.environ.get("PREC", "f16x3"))
img = torch.randn(32, 3, 322, 322, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1234))
for _ in range(int(os.environ.get("CALLS", 1))):
    out = ext(img)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.abs().max()))
