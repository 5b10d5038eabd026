"""BASELINE config 3: 10k x 49152 database, 1k queries, cosine top-5 on one B200 (retrieval micro-benchmark of
SURVEY.md 8(d): DB = normalised randn, queries = DB rows + 0.1*noise, seed 7) -- time and size-independent checks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anyloc_b200 import utilities as u

n_db, n_q, Dv, k = int(os.environ.get("NDB", 10000)), int(os.environ.get("NQ", 1000)), 49152, 5
g = torch.Generator(device="cuda").manual_seed(7)
db = torch.nn.functional.normalize(torch.randn(n_db, Dv, device="cuda", generator=g), dim=1)
src = torch.randint(0, n_db, (n_q,), device="cuda", generator=g)
qu = db[src] + 0.1 * torch.nn.functional.normalize(torch.randn(n_q, Dv, device="cuda", generator=g), dim=1)
for _ in range(2):
    dist, idx = u.top_k_search(db, qu, k)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5):
    dist, idx = u.top_k_search(db, qu, k)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
fl = 2.0 * n_q * n_db * Dv
print(f"c3 retrieval {n_q}x{n_db}x{Dv} top-{k}: {ms:.2f} ms  ({fl/ms/1e9:.1f} TFLOP/s algorithmic, {n_q/ms*1e3:.0f} queries/s)")
assert torch.equal(idx[:, 0], src), "rank-1 must be the source row"
assert bool((dist[:, :-1] >= dist[:, 1:]).all())
# exact check on a slice of queries in fp64
ref = (torch.nn.functional.normalize(qu[:64]).double() @ db.double().T).topk(k, dim=1)
print("top-5 identical to fp64 on 64 queries:", bool(torch.equal(idx[:64], ref.indices)),
      " max rel dist err:", float(((dist[:64].double() - ref.values).abs() / ref.values.abs()).max()))
