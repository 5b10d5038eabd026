"""TEST INFRASTRUCTURE -- a CPU double of the product's DEVICE seams, so that the product's HOST logic (argument
handling, caches, ragged lists, chunking, the exact call sequence of the reference's unmodified driver script) can
be exercised on a box without a GPU.  Every seam that would enqueue a kernel through the C ABI is replaced by the
oracle's arithmetic (oracle/anyloc_oracle.py, oracle/fpk_restated.py); nothing here is reachable from the product
(`anyloc_b200` never imports `tests` or `oracle`), which still has no CPU fallback.

    with cpu_double():
        ...   # anyloc_b200.utilities.{DinoV2ExtractFeatures, VLAD, FlatIndex, get_top_k_recall} run on CPU tensors
"""
import contextlib

import torch
from torch.nn import functional as F

from oracle import anyloc_oracle as ao
from oracle import dinov2_restated as dr
from oracle import fpk_restated as fpk


class _DoubleVit:
    """stands in for anyloc_b200.vit.VitWeights: the restated hub model with the given state_dict"""

    def __init__(self, name, state_dict, device, depth=None, pair="tf32"):
        self.name, self.device, self.pair = name, torch.device("cpu"), pair
        n_blocks = 1 + max(int(k.split(".")[1]) for k in state_dict if k.startswith("blocks."))
        self.depth = n_blocks if depth is None else min(depth, n_blocks)
        with torch.device("meta"):
            self.model = dr.DinoVisionTransformer(name, depth_override=self.depth)
        sd = {k: v.detach().cpu().float() for k, v in state_dict.items()
              if not k.startswith("blocks.") or int(k.split(".")[1]) < self.depth}
        self.model.load_state_dict(sd, strict=False, assign=True)
        self.model.eval()
        self.dim = self.model.embed_dim

    def extract(self, img, layer, facet="value", use_cls=False, norm_descs=True, engine="auto"):
        if img.dim() != 4 or img.shape[1] != 3:
            raise ValueError(f"expected an image batch [B,3,H,W], got {tuple(img.shape)}")
        if img.shape[2] % 14 or img.shape[3] % 14:
            raise ValueError("image size is not a multiple of the patch size 14")
        return ao.extract_features(self.model, img.float(), layer, facet, use_cls, norm_descs)


@contextlib.contextmanager
def cpu_double(state_dict_for=None):
    """`state_dict_for(name)` supplies the weights `resolve_state_dict` would have loaded from the hub."""
    from anyloc_b200 import _lib, utilities as u, vit as _vit
    cpu = torch.device("cpu")
    saved = []

    def patch(obj, name, new):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, new)

    def kmeans_assign(self, x, centers):
        return fpk.KMeans(self.n_clusters, mode=self.mode).max_sim(a=x, b=centers)[1].to(torch.int32)

    def kmeans_update(self, x, labels, c):
        K = c.shape[0]
        mask = (labels.long()[None].expand(K, -1) == torch.arange(K)[:, None]).to(x.dtype)
        new_c = mask @ x / mask.sum(-1)[..., :, None]
        new_c[new_c != new_c] = 0
        return new_c, float((new_c - c).pow(2).sum())

    def vlad_run(self, feats, n_valid, dev, want_labels=False):
        centers = torch.as_tensor(self.c_centers).float()
        outs, extra = [], []
        for b in range(feats.shape[0]):
            x = feats[b] if n_valid is None else feats[b, :int(n_valid[b])]
            if self.vlad_mode == "soft":
                outs.append(ao.vlad_generate_soft(x, centers, self.soft_temp, self.intra_norm, self.norm_descs))
                a = ao.vlad_soft_assign(x, centers, self.soft_temp)
                extra.append(F.pad(a, (0, 0, 0, feats.shape[1] - a.shape[0])))
            else:
                outs.append(ao.vlad_generate(x, centers, self.intra_norm, self.norm_descs, self.mode))
                lab = ao.vlad_labels(x, centers, self.mode).to(torch.int32)
                extra.append(F.pad(lab, (0, feats.shape[1] - lab.shape[0]), value=-1))
        return torch.stack(outs), (torch.stack(extra) if want_labels else None)

    def residuals(self, x, dev):
        xn = F.normalize(x) if self.norm_descs else x
        return xn[:, None, :] - torch.as_tensor(self.c_centers).float()[None, :, :]

    def from_residuals(self, resid, labels, assign, dev):
        N, K, D = resid.shape
        out = torch.zeros(K * D)
        if labels is not None:                              # utilities.py:853-861
            for k in set(labels.tolist()):
                cd = resid[labels == k, k].sum(dim=0)
                out[k * D:(k + 1) * D] = F.normalize(cd, dim=0) if self.intra_norm else cd
        else:                                               # :879-887
            for k in range(K):
                cd = (assign[:, k][:, None, None] * resid).reshape(-1, D).sum(dim=0)
                out[k * D:(k + 1) * D] = F.normalize(cd, dim=0) if self.intra_norm else cd
        return F.normalize(out, dim=0)

    def index_reserve(self, capacity, dev):
        self.capacity, self._dev = capacity, cpu
        self._rows = getattr(self, "_rows", [])

    def index_add(self, x):
        x = torch.as_tensor(x).float()
        if x.shape[1] != self.d:
            raise ValueError(f"index dimension {self.d}, got rows of {x.shape[1]}")
        self._rows = getattr(self, "_rows", []) + [x]
        self.ntotal += x.shape[0]

    def index_rows(self):
        if getattr(self, "_slab", None) is not None:             # rows placed with add_at()
            return self._slab[:self.ntotal]
        return torch.cat(self._rows)

    def index_search(self, qu, k, n_q_chunk=4096):
        if self.ntotal == 0:
            raise ValueError("search on an empty index")
        return ao.top_k(index_rows(self), torch.as_tensor(qu).float(), k, self.method, self.norm_descs)

    def index_add_at(self, x, row_offset):
        x = torch.as_tensor(x).float()
        if row_offset < 0 or row_offset + x.shape[0] > self.capacity:
            raise ValueError("rows outside the reserved capacity")
        if getattr(self, "_slab", None) is None:
            self._slab = torch.zeros(self.capacity, self.d)
        self._slab[row_offset:row_offset + x.shape[0]] = x
        self.ntotal = max(self.ntotal, row_offset + x.shape[0])

    def index_reset(self):
        self.ntotal, self._rows, self._slab = 0, [], None

    patch(_lib, "require_cuda", lambda device=None: cpu)
    patch(_vit, "VitWeights", _DoubleVit)
    if state_dict_for is not None:
        patch(_vit, "resolve_state_dict", lambda name, device: state_dict_for(name))
    patch(u, "_normalize_rows_dev", lambda x: F.normalize(x))
    patch(u._KMeans, "_assign", kmeans_assign)
    patch(u._KMeans, "_update", kmeans_update)
    patch(u.VLAD, "_run", vlad_run)
    patch(u.VLAD, "_residuals_dev", residuals)
    patch(u.VLAD, "_from_residuals_dev", from_residuals)
    patch(u.FlatIndex, "_reserve", index_reserve)
    patch(u.FlatIndex, "add", index_add)
    patch(u.FlatIndex, "search", index_search)
    patch(u.FlatIndex, "add_at", index_add_at)
    patch(u.FlatIndex, "reset", index_reset)
    try:
        yield
    finally:
        for obj, name, old in reversed(saved):
            setattr(obj, name, old)
