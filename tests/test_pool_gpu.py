"""GPU parity of the sibling aggregators (anyloc_pool through utilities.pool_descriptors) against the oracle's
restatement of scripts/dino_v2_gem.py:170-189 and scripts/dino_v2_gp.py:130-135 (fp64), tolerance 1e-4."""
import pytest
import torch

from oracle import anyloc_oracle as ao
from tests.util import rel_inf

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def u(cuda):
    from anyloc_b200 import utilities
    return utilities


@pytest.mark.parametrize("B,N,D", [(3, 529, 1536), (2, 1369, 1024), (4, 5, 36), (1, 1, 384)])
def test_pool_modes(u, B, N, D):
    x = torch.nn.functional.normalize(torch.randn(B, N, D, generator=torch.Generator().manual_seed(N + D)), dim=-1)
    for method in ("average", "max"):
        out = u.pool_descriptors(x, method)
        assert out.device.type == "cpu" and out.shape == (B, D)
        ref = ao.pool_descriptors(x.double(), method)
        if method == "max":
            assert torch.equal(out, ref.float())
        else:
            assert rel_inf(out, ref) < TOL
    for p, use_abs in ((3, False), (3, True), (2, False), (2.5, True), (5, False)):
        out = u.pool_descriptors(x.cuda(), "gem", gem_p=p, gem_use_abs=use_abs)
        assert out.is_cuda
        ref = ao.gem_descriptors(x.double(), p, use_abs)
        assert rel_inf(out.cpu(), ref) < TOL


def test_pool_edge_cases(u):
    x = torch.randn(2, 9, 8)
    x[0, 3, 2] = float("nan")
    out = u.pool_descriptors(x, "max")
    assert torch.isnan(out[0, 2]) and torch.equal(out[1], x[1].max(dim=0)[0])
    z = torch.zeros(1, 4, 8)
    assert torch.equal(u.pool_descriptors(z, "gem"), torch.zeros(1, 8))      # sign(0) * 0 = 0
    with pytest.raises(NotImplementedError):
        u.pool_descriptors(x, "median")
    with pytest.raises(AssertionError):
        u.pool_descriptors(torch.randn(4, 8), "gem")
