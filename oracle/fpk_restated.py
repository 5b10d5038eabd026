"""TEST INFRASTRUCTURE -- restatement of fast-pytorch-kmeans==0.1.6 (not vendored
by the reference; pinned in /root/reference/setup_conda.sh:224,
requirements.txt:54).  Call sites in the reference: utilities.py:766 (ctor),
:772 (centroids=), :786-787 (fit, centroids), :849 (predict).

Published algorithm:
  cos_sim(a,b) = a/(|a|+1e-8) @ (b/(|b|+1e-8)).T
  euc_sim(a,b) = 2 a@b.T - |a|^2[:,None] - |b|^2[None,:]
  max_sim      = sim.max(dim=-1)            (first index on exact ties, CPU)
  fit_predict  = random-choice init (numpy RNG) + <=max_iter Lloyd steps,
                 centroids = plain means of members (NaN -> 0 for empty), stop
                 when sum((c_new-c_old)^2) <= tol.
PARITY UNPINNED at this boundary (no reference-held vectors exist).
"""
import numpy as np
import torch


class KMeans:
    def __init__(self, n_clusters, max_iter=100, tol=0.0001, verbose=0,
                 mode="euclidean", minibatch=None):
        self.n_clusters = n_clusters
        self.max_iter = max_iter
        self.tol = tol
        self.verbose = verbose
        self.mode = mode
        self.minibatch = minibatch
        self.centroids = None

    @staticmethod
    def cos_sim(a, b):
        a_norm = a.norm(dim=-1, keepdim=True)
        b_norm = b.norm(dim=-1, keepdim=True)
        a = a / (a_norm + 1e-8)
        b = b / (b_norm + 1e-8)
        return a @ b.transpose(-2, -1)

    @staticmethod
    def euc_sim(a, b):
        return 2 * a @ b.transpose(-2, -1) - (a ** 2).sum(dim=1)[..., :, None] \
            - (b ** 2).sum(dim=1)[..., None, :]

    def max_sim(self, a, b):
        if self.mode == "cosine":
            sim = self.cos_sim(a, b)
        elif self.mode == "euclidean":
            sim = self.euc_sim(a, b)
        else:
            raise NotImplementedError(self.mode)
        return sim.max(dim=-1)

    def predict(self, X):
        return self.max_sim(a=X, b=self.centroids)[1]

    def fit_predict(self, X, centroids=None):
        n = X.shape[0]
        if centroids is None:
            self.centroids = X[np.random.choice(n, size=[self.n_clusters], replace=False)]
        else:
            self.centroids = centroids
        closest = None
        for _ in range(self.max_iter):
            closest = self.max_sim(a=X, b=self.centroids)[1]
            expanded = closest[None].expand(self.n_clusters, -1)
            mask = (expanded == torch.arange(self.n_clusters, device=X.device)[:, None]).to(X.dtype)
            c_grad = mask @ X / mask.sum(-1)[..., :, None]
            c_grad[c_grad != c_grad] = 0
            error = (c_grad - self.centroids).pow(2).sum()
            self.centroids = c_grad
            if error <= self.tol:
                break
        return closest

    def fit(self, X, centroids=None):
        self.fit_predict(X, centroids)
