"""VLAD-only driver for ncu / timing: c2 (B=32,N=529,D=1536,K=32) and c5 (B=64,N=1369,D=1024,K=128) shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anyloc_b200 import utilities as u

for (B, N, D, K) in [(32, 529, 1536, 32), (64, 1369, 1024, 128)]:
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.nn.functional.normalize(torch.randn(B, N, D, device="cuda", generator=g), dim=-1)
    c = 0.7 * x.reshape(-1, D)[torch.randperm(B * N, device="cuda", generator=g)[:K]].contiguous()
    v = u.VLAD(K); v.kmeans = u._KMeans(K, mode="cosine"); v.c_centers = v.kmeans.centroids = c; v.desc_dim = D
    for _ in range(3):
        out = v.generate_multi(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        out = v.generate_multi(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    by = 4.0 * (B * N * D + B * K * D + K * D)
    print(f"B={B} N={N} D={D} K={K}: {ms*1e3:.1f} us per call, {by/ms/1e6:.0f} GB/s algorithmic ({by/1e6:.1f} MB)")
