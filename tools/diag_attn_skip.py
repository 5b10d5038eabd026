"""Where does the tail-skipping attention variant (ANYLOC_ATTN_SKIP=1) go wrong?  One subprocess per shape (a hang
only costs its own timeout); prints the error of the f16 tcgen05 attention against fp64 per 32-row block of the query axis
and per 16-column block of the head dimension, plus the positions of non-finite outputs.
    ANYLOC_ATTN_SKIP=1 python tools/diag_attn_skip.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(1, 64, 1), (1, 65, 1), (1, 129, 1), (2, 257, 2), (1, 530, 2), (2, 530, 24), (1, 1370, 2)]


def one(B, T, heads):
    sys.path.insert(0, ROOT)
    import torch
    from anyloc_b200 import _lib as L
    lib = L.load()
    D = heads * 64
    g = torch.Generator(device="cuda").manual_seed(T)
    qkv = torch.randn(B, T, 3 * D, device="cuda", generator=g) * 1.5
    hi, lo = torch.empty_like(qkv), torch.empty_like(qkv)
    L.check(lib.anyloc_split_tf32(L.ptr(qkv), L.ptr(hi), L.ptr(lo), qkv.numel(), L.stream_ptr()), "split")
    oh = torch.zeros(B, T, D, device="cuda", dtype=torch.float16)
    ol = torch.zeros(B, T, D, device="cuda", dtype=torch.float16)
    L.check(lib.anyloc_attention(L.ptr(hi), L.ptr(lo), B, T, D, heads, L.ptr(oh), L.ptr(ol), L.PAIR["f16"],
                                 L.ENGINE["tc3"], L.stream_ptr()), "attn")
    torch.cuda.synchronize()
    out = (oh.double() + ol.double()) / L.ACT_SCALE
    q, k, v = (t.reshape(B, T, heads, 64).transpose(1, 2).double() for t in qkv.chunk(3, dim=-1))
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v).transpose(1, 2).reshape(B, T, D)
    bad = ~torch.isfinite(out)
    err = (out - ref).abs()
    err[bad] = float("inf")
    scale = float(ref.abs().max())
    print(f"B={B} T={T} heads={heads}: rel err {float(err.max()) / scale:.3e}, non-finite {int(bad.sum())} of {out.numel()}")
    if float(err.max()) / scale > 1e-5:
        rows = err.amax(dim=(0, 2))                       # per query row
        blocks = [float(rows[i:i + 32].max()) / scale for i in range(0, T, 32)]
        print("   per 32-row block:", " ".join(f"{b:.1e}" for b in blocks))
        cols = err.reshape(B, T, heads, 4, 16).amax(dim=(0, 1, 2, 4))
        print("   per 16-dim part :", " ".join(f"{float(c) / scale:.1e}" for c in cols))
        per_b = err.amax(dim=(1, 2))
        print("   per image       :", " ".join(f"{float(c) / scale:.1e}" for c in per_b))


if __name__ == "__main__":
    if len(sys.argv) == 4:
        one(*map(int, sys.argv[1:]))
    else:
        for B, T, h in SHAPES:
            try:
                r = subprocess.run([sys.executable, __file__, str(B), str(T), str(h)], capture_output=True, text=True,
                                   timeout=40)
                print(r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
            except subprocess.TimeoutExpired:
                print(f"B={B} T={T} heads={h}: TIMEOUT (hang)", flush=True)
