import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anyloc_b200 import _lib as L
lib = L.load()
def split(x):
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    L.check(lib.anyloc_split_tf32(L.ptr(x), L.ptr(hi), L.ptr(lo), x.numel(), L.stream_ptr()), "split"); return hi, lo
fn = lib.anyloc_attention_tc16_debug
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
B, T, heads = 32, 530, 24
D = heads * 64
qkv = torch.randn(B, T, 3 * D, device="cuda") * 1.5
qh, ql = split(qkv)
oh, ol = torch.zeros(B, T, D, device="cuda", dtype=torch.float16), torch.zeros(B, T, D, device="cuda", dtype=torch.float16)
dbg = torch.full((256,), float("nan"), device="cuda")
rc = fn(L.ptr(qh), L.ptr(ql), B, T, D, heads, L.ptr(oh), L.ptr(ol), L.ptr(dbg), L.stream_ptr())
torch.cuda.synchronize()
ts = dbg.reshape(16, 16).cpu()
print("rc", rc, "| softmax [before s_full, after s_full, after ld, after max xchg, after exp, after o_full, after P st, after arrive] | MMA [before S issue, after S issue, after p_full, after PV issue]")
for j in range((T + 127) // 128):
    print(j, [int(v) if v == v else -1 for v in ts[j, :8].tolist()], "|", [int(v) if v == v else -1 for v in ts[j, 8:12].tolist()])
