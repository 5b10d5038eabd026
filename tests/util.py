"""Shared helpers for the parity tests (the oracle is the checker, never the thing under test)."""
import ast
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_cases(fname):
    z = np.load(os.path.join(GOLDEN, fname), allow_pickle=False)
    cases = {}
    for key in z.files:
        if "/" in key:
            name, field = key.split("/", 1)
            cases.setdefault(name, {})[field] = z[key]
        else:
            cases.setdefault("", {})[key] = z[key]
    return cases


def case_kwargs(case):
    return ast.literal_eval(str(case["kw"])) if "kw" in case else {}


def rel_inf(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def make_vlad(u, K, centers, **kw):
    """product VLAD object with a given vocabulary (what `fit` from a c_centers.pt cache yields)."""
    v = u.VLAD(K, **kw)
    v.kmeans = u._KMeans(K, mode=v.mode)
    v.kmeans.centroids = torch.as_tensor(centers)
    v.c_centers = torch.as_tensor(centers)
    v.desc_dim = centers.shape[1]
    return v
