"""TEST INFRASTRUCTURE -- restatement of the two faiss==1.7.2 indices the
reference uses (not vendored; pinned in /root/reference/setup_conda.sh:222).
Call sites: utilities.py:440 (IndexFlatIP), :442 (IndexFlatL2), :446-447 (GPU
resources), :449 (add), :450 (search).

Published behaviour: exact brute-force search.  IndexFlatIP.search returns the
k largest inner products sorted descending; IndexFlatL2.search the k smallest
SQUARED L2 distances sorted ascending.  faiss' order among exactly-equal scores
is heap dependent and unpinned; this restatement (and the product) define
lowest-database-index-first.  With faiss.contrib.torch_utils (utilities.py:14)
torch inputs give torch outputs (float32 distances, int64 indices).
PARITY UNPINNED at this boundary (no reference-held vectors exist).
"""
import numpy as np
import torch


def _stable_topk(score: torch.Tensor, k: int, largest: bool):
    # stable sort => lowest index first among equal scores
    key = -score if largest else score
    order = torch.sort(key, dim=1, stable=True)[1][:, :k]
    return torch.gather(score, 1, order), order


class _IndexFlat:
    _largest = True

    def __init__(self, d):
        self.d = d
        self._db = None
        self._numpy = False

    @property
    def ntotal(self):
        return 0 if self._db is None else self._db.shape[0]

    def add(self, x):
        if isinstance(x, np.ndarray):
            self._numpy = True
            x = torch.from_numpy(x)
        x = x.detach().to(torch.float32).cpu()
        assert x.shape[1] == self.d
        self._db = x if self._db is None else torch.cat([self._db, x])

    def _score(self, q):
        raise NotImplementedError

    def search(self, q, k):
        as_numpy = isinstance(q, np.ndarray)
        if as_numpy:
            q = torch.from_numpy(q)
        q = q.detach().to(torch.float32).cpu()
        dist, idx = _stable_topk(self._score(q), k, self._largest)
        idx = idx.to(torch.int64)
        if as_numpy:
            return dist.numpy(), idx.numpy()
        return dist, idx


class IndexFlatIP(_IndexFlat):
    _largest = True

    def _score(self, q):
        return q @ self._db.T


class IndexFlatL2(_IndexFlat):
    _largest = False

    def _score(self, q):
        # exact squared L2, the quantity IndexFlatL2 reports
        return (q * q).sum(1)[:, None] - 2.0 * (q @ self._db.T) + (self._db * self._db).sum(1)[None, :]


class StandardGpuResources:
    pass


def index_cpu_to_gpu(res, dev, index):
    return index
