"""TEST INFRASTRUCTURE -- CPU restatement (torch, fp32 with optional fp64) of the
reference's hot path.  Each function cites the reference lines it follows
(/root/reference/utilities.py).  Pinned against the verbatim reference import
through tests/golden/*.npz (see tests/golden/make_golden.py).  Travels to the
GPU box (the reference tree does not).
"""
import numpy as np
import torch
from torch.nn import functional as F


# ---------------------------------------------------------------- extractor
def extract_features(model, img, layer, facet="value", use_cls=False, norm_descs=True):
    """utilities.py:263-285 with the hook of :245-252 evaluated explicitly:
    facet 'token'   = output of blocks[layer];
    facet q/k/v     = thirds of blocks[layer].attn.qkv(norm1(x_layer)).
    Early exit after the hooked module is output-identical to the full forward the
    reference runs (its result is discarded, utilities.py:269-273)."""
    with torch.no_grad():
        x = model.prepare_tokens(img)
        for blk in model.blocks[:layer]:
            x = blk(x)
        blk = model.blocks[layer]
        if facet == "token":
            res = blk(x)
        else:
            res = blk.attn.qkv(blk.norm1(x))
        if not use_cls:
            res = res[:, 1:, ...]
        if facet in ("query", "key", "value"):
            d = res.shape[2] // 3
            i = {"query": 0, "key": 1, "value": 2}[facet]
            res = res[:, :, i * d:(i + 1) * d]
    if norm_descs:
        res = F.normalize(res, dim=-1)
    return res


def extract_features_full_forward(model, img, layer, facet="value", use_cls=False, norm_descs=True):
    """The reference's execution strategy (for the CPU-baseline timing): run the WHOLE model with a
    forward hook on blocks[layer] / blocks[layer].attn.qkv and discard the model output
    (utilities.py:245-252, :268-273).  Output-identical to extract_features()."""
    captured = {}
    mod = model.blocks[layer] if facet == "token" else model.blocks[layer].attn.qkv
    handle = mod.register_forward_hook(lambda m, i, o: captured.__setitem__("out", o))
    try:
        with torch.no_grad():
            model(img)
            res = captured["out"]
            if not use_cls:
                res = res[:, 1:, ...]
            if facet in ("query", "key", "value"):
                d = res.shape[2] // 3
                i = {"query": 0, "key": 1, "value": 2}[facet]
                res = res[:, :, i * d:(i + 1) * d]
    finally:
        handle.remove()
    if norm_descs:
        res = F.normalize(res, dim=-1)
    return res


# --------------------------------------------------------------------- pre-processing
def preprocess(img_u8, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), patch=14, resize=None,
               interpolation="bilinear"):
    """`base_transform` (dvgl_benchmark/datasets_ws.py:20-23: ToTensor = HWC uint8 -> CHW float / 255, Normalize =
    (x - mean) / std), optionally `T.functional.resize(img, resize)` (:233-235; on a float tensor torchvision calls
    torch interpolate(mode, align_corners=False, antialias=True) -- `interpolation="bicubic"` is the demo's
    down-scaling, demo/anyloc_vlad_generate.py:165-177), then `T.CenterCrop(((h // 14) * 14, (w // 14) * 14))`
    (scripts/dino_v2_vlad.py:174-176; torchvision puts the window at int(round((h - h_new) / 2.0))).
    img_u8 [H,W,3] uint8 -> [3,h_new,w_new]."""
    x = torch.as_tensor(img_u8).permute(2, 0, 1).to(torch.float32).div(255)
    x = (x - torch.tensor(mean, dtype=torch.float32)[:, None, None]) / torch.tensor(std, dtype=torch.float32)[:, None, None]
    if resize is not None:
        x = F.interpolate(x[None], size=tuple(resize), mode=interpolation, align_corners=False, antialias=True)[0]
    h, w = x.shape[1:]
    hn, wn = (h // patch) * patch, (w // patch) * patch
    top, left = int(round((h - hn) / 2.0)), int(round((w - wn) / 2.0))
    return x[:, top:top + hn, left:left + wn]


# --------------------------------------------------------------------- VLAD
def assign_similarity(x, centers, dist_mode="cosine"):
    """fpk.KMeans.max_sim as reached from utilities.py:849 (`predict` on the
    descriptors as passed in, i.e. NOT re-normalised in that scope)."""
    if dist_mode == "cosine":
        a = x / (x.norm(dim=-1, keepdim=True) + 1e-8)
        b = centers / (centers.norm(dim=-1, keepdim=True) + 1e-8)
        return a @ b.T
    if dist_mode == "euclidean":
        return 2 * x @ centers.T - (x ** 2).sum(1)[:, None] - (centers ** 2).sum(1)[None, :]
    raise NotImplementedError(dist_mode)


def vlad_labels(x, centers, dist_mode="cosine"):
    return assign_similarity(x, centers, dist_mode).max(dim=-1)[1]


def vlad_generate(x, centers, intra_norm=True, norm_descs=True, dist_mode="cosine",
                  labels=None, dtype=None):
    """utilities.py:819-890 hard branch (:841-861) + residuals (:956-962), without
    materialising [N,K,D].  `labels` may be forced (margin-aware parity tests);
    `dtype=torch.float64` gives the high-precision variant."""
    x = torch.as_tensor(x)
    centers = torch.as_tensor(centers)
    if dtype is not None:
        x, centers = x.to(dtype), centers.to(dtype)
    K, D = centers.shape
    if labels is None:
        labels = vlad_labels(x, centers, dist_mode)          # :849
    xn = F.normalize(x) if norm_descs else x                  # :959-960
    out = torch.zeros(K * D, dtype=x.dtype)                   # :840
    for k in sorted(set(labels.tolist())):                    # :854-855
        cd = (xn[labels == k] - centers[k][None]).sum(0)      # :858 (residual :961-962)
        if intra_norm:
            cd = F.normalize(cd, dim=0)                       # :859-860
        out[k * D:(k + 1) * D] = cd                           # :861
    return F.normalize(out, dim=0)                            # :889


def vlad_generate_faithful(x, centers, intra_norm=True, norm_descs=True, dist_mode="cosine"):
    """Same result as vlad_generate(), with the reference's memory behaviour (for CPU-baseline timing):
    the full [N,K,D] residual tensor (utilities.py:961-962) and boolean-mask gathers (:858)."""
    K, D = centers.shape
    xn = F.normalize(x) if norm_descs else x
    residuals = xn[:, None, :] - centers[None, :, :]
    labels = vlad_labels(x, centers, dist_mode)
    out = torch.zeros(K * D)
    for k in set(labels.numpy()):
        cd = residuals[labels == k, k].sum(dim=0)
        if intra_norm:
            cd = F.normalize(cd, dim=0)
        out[k * D:(k + 1) * D] = cd
    return F.normalize(out, dim=0)


def vlad_soft_assign(x, centers, soft_temp=1.0):
    """utilities.py:870-875: softmax(soft_temp * F.cosine_similarity(x, c)) on the descriptors as passed."""
    cos = F.cosine_similarity(x[:, None, :], centers[None, :, :], dim=2)
    return F.softmax(soft_temp * cos, dim=1)


def vlad_generate_soft(x, centers, soft_temp=1.0, intra_norm=True, norm_descs=True, dtype=None):
    """utilities.py:862-887 soft branch.  For cluster k the reference sums w_k * residuals over BOTH q and c
    (:881-884), i.e. V_k = sum_q a[q,k] * sum_c (x^_q - c_c); restated with the same [N,K,D] residual tensor
    and the same flattened summation so that fp32 results are bit-identical."""
    x = torch.as_tensor(x)
    centers = torch.as_tensor(centers)
    if dtype is not None:
        x, centers = x.to(dtype), centers.to(dtype)
    K, D = centers.shape
    xn = F.normalize(x) if norm_descs else x                  # :959-960
    residuals = xn[:, None, :] - centers[None, :, :]          # :961-962
    a = vlad_soft_assign(x, centers, soft_temp)               # :870-875
    out = torch.zeros(K * D, dtype=x.dtype)
    for k in range(K):                                        # :880
        cd = (a[:, k][:, None, None] * residuals).reshape(-1, D).sum(dim=0)   # :881-884
        if intra_norm:
            cd = F.normalize(cd, dim=0)                       # :885-886
        out[k * D:(k + 1) * D] = cd
    return F.normalize(out, dim=0)                            # :889


def vlad_generate_soft_closed(x, centers, soft_temp=1.0, intra_norm=True, norm_descs=True, dtype=torch.float64):
    """The soft branch in closed form, V_k = K sum_q a_qk x^_q - (sum_q a_qk) sum_c c_c: equal to
    vlad_generate_soft() up to summation order (tests/test_oracle_cpu.py checks that on the golden cases) without
    the [N,K,D] tensor -- for pipeline-sized checks."""
    x, centers = torch.as_tensor(x).to(dtype), torch.as_tensor(centers).to(dtype)
    K, D = centers.shape
    a = vlad_soft_assign(x, centers, soft_temp)
    xn = F.normalize(x) if norm_descs else x
    v = K * (a.T @ xn) - a.sum(0)[:, None] * centers.sum(0)[None]
    if intra_norm:
        v = F.normalize(v, dim=1)
    return F.normalize(v.reshape(-1), dim=0)


def vlad_generate_multi(xs, centers, **kw):
    """utilities.py:892-926."""
    return torch.stack([vlad_generate(x, centers, **kw) for x in xs])


def label_margins(x, centers, dist_mode="cosine"):
    """fp64 top1-top2 similarity gap per descriptor and fp64 labels."""
    s = assign_similarity(torch.as_tensor(x).double(), torch.as_tensor(centers).double(), dist_mode)
    if s.shape[1] == 1:
        return torch.full((s.shape[0],), float("inf"), dtype=torch.float64), torch.zeros(s.shape[0], dtype=torch.long)
    top2 = s.topk(2, dim=1)[0]
    return top2[:, 0] - top2[:, 1], s.max(dim=1)[1]


# ---------------------------------------------------------------- sibling aggregators
def gem_descriptors(patch_descs, gem_p=3, gem_use_abs=False):
    """scripts/dino_v2_gem.py:170-189 (`get_gem_descriptors`; the script is argparse-driven and cannot be
    imported, so this restatement is PARITY UNPINNED -- it follows the cited lines)."""
    if gem_use_abs:
        return torch.mean(torch.abs(patch_descs) ** gem_p, dim=-2) ** (1 / gem_p)          # :173-175
    x = torch.mean(patch_descs ** gem_p, dim=-2)                                            # :185
    g = x.to(torch.complex128 if x.dtype == torch.float64 else torch.complex64) ** (1 / gem_p)   # :186
    return torch.abs(g) * torch.sign(x)                                                     # :187


def pool_descriptors(patch_descs, method):
    """scripts/dino_v2_gp.py:130-135 (PARITY UNPINNED, same reason)."""
    if method == "average":
        return torch.mean(patch_descs, dim=1)
    if method == "max":
        return torch.max(patch_descs, dim=1)[0]
    raise NotImplementedError(f"ID: {method}")


# ---------------------------------------------------------------- retrieval
def top_k(db, qu, k, method="cosine", norm_descs=True, dtype=None):
    """utilities.py:433-450 (normalise, exact IP / squared-L2 search, k best, sorted,
    lowest index first among equals)."""
    db, qu = torch.as_tensor(db), torch.as_tensor(qu)
    if dtype is not None:
        db, qu = db.to(dtype), qu.to(dtype)
    if qu.dim() == 1:
        qu = qu.unsqueeze(0)
    if norm_descs:
        db, qu = F.normalize(db), F.normalize(qu)
    if method == "cosine":
        score, largest = qu @ db.T, True
    elif method == "l2":
        score = (qu * qu).sum(1)[:, None] - 2.0 * (qu @ db.T) + (db * db).sum(1)[None, :]
        largest = False
    else:
        raise NotImplementedError(f"Method: {method}")
    order = torch.sort(-score if largest else score, dim=1, stable=True)[1][:, :k]
    return torch.gather(score, 1, order), order


def recalls_from_indices(indices, top_k_vals, gt_pos, use_percentage=True,
                         sub_sample_db=1, sub_sample_qu=1):
    """utilities.py:451-468."""
    indices = np.asarray(indices)
    recalls = dict(zip(top_k_vals, [0] * len(top_k_vals)))
    for i_qu, qu_retr in enumerate(indices):
        for i_rec in top_k_vals:
            correct = gt_pos[i_qu * sub_sample_qu]
            if np.any(np.isin(qu_retr[:i_rec] * sub_sample_db, correct)):
                recalls[i_rec] += 1
    if use_percentage:
        for k in recalls:
            recalls[k] /= len(indices)
    return recalls


def get_top_k_recall(top_k_vals, db, qu, gt_pos, method="cosine", norm_descs=True,
                     use_percentage=True, sub_sample_db=1, sub_sample_qu=1):
    dist, idx = top_k(db, qu, max(top_k_vals), method, norm_descs)
    return dist, idx, recalls_from_indices(idx, top_k_vals, gt_pos, use_percentage,
                                           sub_sample_db, sub_sample_qu)


# ------------------------------------------------------------ synthetic data
def clustered_features(n, d, k, seed=0, kappa_noise=0.35, centre_norm=0.8):
    """K well separated lobes on the sphere (large assignment margins) + the
    vocabulary that generated them (un-normalised centres, norm<1 like k-means
    means of unit vectors)."""
    g = torch.Generator().manual_seed(seed)
    mu = F.normalize(torch.randn(k, d, generator=g), dim=1)
    lab = torch.randint(0, k, (n,), generator=g)
    x = F.normalize(mu[lab] + kappa_noise / d ** 0.5 * torch.randn(n, d, generator=g), dim=1)
    centres = centre_norm * mu * (1.0 + 0.1 * torch.rand(k, 1, generator=g))
    return x, centres, lab


# ---------------------------------------------------------------- PCA (descriptor dimensionality reduction)
def reduce_pca(train_descs, test_descs, lower_dim, low_factor=0.0, fallback=256, svd_solver="full", whitening=False):
    """utilities.py:522-586, restated on the same sklearn calls (sklearn.decomposition.PCA; pinned against the verbatim
    import by tests/test_oracle_cpu.py::test_reduce_pca_matches_reference)."""
    from sklearn.decomposition import PCA
    assert 0 <= low_factor <= 1
    if low_factor == 0.0:                                                    # :560-563
        pca = PCA(lower_dim, svd_solver=svd_solver, whiten=whitening)
        return pca.fit_transform(train_descs), pca.transform(test_descs)
    n_samples, n_components = train_descs.shape                              # :565
    if n_samples < n_components:                                             # :566-574
        both = np.concatenate((train_descs.copy(), test_descs.copy()))
        both = PCA(fallback, svd_solver=svd_solver).fit_transform(both)
        train_descs, test_descs = both[:n_samples], both[n_samples:]
    n_low = int(low_factor * lower_dim)                                      # :575-576
    n_top = lower_dim - n_low
    pca = PCA(train_descs.shape[1], svd_solver=svd_solver)                   # :578-580
    pca.fit(train_descs)
    basis = np.concatenate((pca.components_[:n_top], pca.components_[-n_low:]))   # :581-582
    return (train_descs - pca.mean_) @ basis.T, (test_descs - pca.mean_) @ basis.T   # :583-584
