"""GPU parity of reduce_pca (utilities.py:522-586; scripts/dino_v2_vlad.py:357-369 reduces the VLADs with it): the
device PCA (fp64 eigen-decomposition of the Gram / covariance matrix, projections as tcgen05 GEMMs) against the
reference's own sklearn calls (oracle restatement, pinned to the verbatim import on the CPU tier).  Data with a
geometrically decaying spectrum, so that every retained component is well separated (PCA directions of near-equal
singular values are not determined by the data; sklearn itself solves in fp32)."""
import numpy as np
import pytest
import torch

from oracle import anyloc_oracle as ao
from tests.util import rel_inf

pytestmark = pytest.mark.gpu


def spectrum_data(n, d, rank, decay, seed, n_test=37):
    g = np.random.default_rng(seed)
    basis = np.linalg.qr(g.standard_normal((d, rank)))[0]                  # [d, rank] orthonormal
    scales = decay ** np.arange(rank)
    offset = 0.3 * g.standard_normal(d)                                    # non-zero mean: centring matters

    def make(m):
        return ((g.standard_normal((m, rank)) * scales) @ basis.T + offset).astype(np.float32)
    return make(n), make(n_test)


@pytest.fixture(scope="module")
def u(cuda):
    from anyloc_b200 import utilities
    return utilities


@pytest.mark.parametrize("n,d,k,whiten", [(300, 96, 16, False), (300, 96, 16, True), (120, 512, 24, True),
                                          (1000, 64, 32, False)])
def test_reduce_pca_direct(u, n, d, k, whiten):
    tr, te = spectrum_data(n, d, min(n, d, 48), 0.88, seed=n + d)
    r_tr, r_te = ao.reduce_pca(tr, te, k, whitening=whiten)
    o_tr, o_te = u.reduce_pca(tr, te, k, whitening=whiten)
    assert type(o_tr) == np.ndarray and o_tr.shape == r_tr.shape and o_te.shape == r_te.shape
    assert o_tr.dtype == np.float32
    assert rel_inf(o_tr, r_tr) < 1e-4 and rel_inf(o_te, r_te) < 1e-4
    # the script L2-normalises what comes back (scripts/dino_v2_vlad.py:358-360, :367-368): same unit rows
    nrm = lambda x: x / np.linalg.norm(x, axis=-1, keepdims=True)
    assert rel_inf(nrm(o_te), nrm(r_te)) < 1e-4


def test_reduce_pca_low_factor(u):
    tr, te = spectrum_data(400, 40, 40, 0.9, seed=9)
    r_tr, r_te = ao.reduce_pca(tr, te, 10, low_factor=0.3)
    o_tr, o_te = u.reduce_pca(tr, te, 10, low_factor=0.3)
    assert o_tr.shape == r_tr.shape == (400, 10)
    assert rel_inf(o_tr, r_tr) < 1e-4 and rel_inf(o_te, r_te) < 1e-4


def test_reduce_pca_torch_inputs_and_errors(u):
    tr, te = spectrum_data(64, 32, 32, 0.8, seed=2)
    o_tr, o_te = u.reduce_pca(torch.from_numpy(tr), torch.from_numpy(te), 8)
    assert isinstance(o_tr, torch.Tensor) and not o_tr.is_cuda
    with pytest.raises(ValueError):
        u.reduce_pca(tr, te, 65)                    # more components than min(n_samples, n_features), like sklearn
