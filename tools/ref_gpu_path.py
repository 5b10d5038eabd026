"""The north star's 10x denominator: 'the reference GPU PyTorch path' on this B200 -- exactly the loop of
scripts/dino_v2_vlad.py:164-188,233-237: restated hub model .cuda() in fp32, ONE image per forward (all blocks + hook),
ret.cpu() per image, then CPU VLAD.generate per image (oracle, [N,K,D] residuals).  Baseline only (torch/cuBLAS
kernels, none of this repo's).  Also: this repo's own drop-in path at batch 1 for comparison."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import anyloc_oracle as ao, dinov2_restated as dr
import bench

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.set_num_threads(bench.usable_cores())
wl = bench.WORKLOADS["c2"]
n_img = int(os.environ.get("NIMG", 16))
model = dr.build(wl["model"], seed=0).cuda()
g = torch.Generator().manual_seed(1234)
imgs = torch.randn(n_img, 3, wl["H"], wl["W"], generator=g)
centers = 0.6 * torch.nn.functional.normalize(torch.randn(wl["K"], model.embed_dim, generator=g), dim=1)

def ref_path(n):
    feats = []
    for i in range(n):
        x = imgs[i:i + 1].cuda()
        feats.append(ao.extract_features_full_forward(model, x, wl["layer"], wl["facet"]).cpu())
    feats = torch.cat(feats)
    return torch.stack([ao.vlad_generate_faithful(f, centers) for f in feats])

ref_path(2)
torch.cuda.synchronize(); t0 = time.perf_counter()
out_ref = ref_path(n_img)
torch.cuda.synchronize(); t_all = time.perf_counter() - t0
t0 = time.perf_counter()
for i in range(n_img):
    ao.extract_features_full_forward(model, imgs[i:i + 1].cuda(), wl["layer"], wl["facet"]).cpu()
t_vit = time.perf_counter() - t0
print(f"reference GPU PyTorch path (fp32, batch 1, CPU VLAD on {bench.usable_cores()} cores): {n_img / t_all:.2f} img/s "
      f"({t_all / n_img * 1e3:.1f} ms/img; ViT part {t_vit / n_img * 1e3:.1f} ms/img)")

# this repo, same calling pattern (batch 1 per call, host tensors into VLAD.generate_multi)
from anyloc_b200 import utilities as u
sd = {k: v for k, v in model.state_dict().items()}
for prec in ("f16x3", "tf32x3"):
    ext = u.DinoV2ExtractFeatures(wl["model"], wl["layer"], wl["facet"], device="cuda", weights=sd, precision=prec)
    vl = u.VLAD(wl["K"]); vl.kmeans = u._KMeans(wl["K"], mode="cosine"); vl.c_centers = vl.kmeans.centroids = centers; vl.desc_dim = model.embed_dim
    def ours(n):
        fs = [ext(imgs[i:i + 1].cuda()).cpu() for i in range(n)]
        return vl.generate_multi(torch.cat(fs))
    ours(2)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = ours(n_img)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    err = float((out - out_ref).abs().max() / out_ref.abs().max())
    print(f"anyloc_b200 drop-in path, batch 1 per call, {prec}: {n_img / t:.2f} img/s ({t / n_img * 1e3:.1f} ms/img); "
          f"descriptor rel err vs reference path {err:.2e}")
    del ext
