"""GPU parity of the fused pre-processing kernel (anyloc_preprocess_u8 through utilities.preprocess_images) against
torchvision's own outputs (tests/golden/preprocess.npz) and the oracle restatement: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import anyloc_oracle as ao
from tests.util import load_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def u(cuda):
    from anyloc_b200 import utilities
    return utilities


@pytest.mark.parametrize("name", sorted(load_cases("preprocess.npz")))
def test_preprocess_golden(u, name):
    c = load_cases("preprocess.npz")[name]
    out = u.preprocess_images(c["img"])
    assert out.is_cuda and out.shape == (1,) + c["out"].shape
    assert torch.equal(out[0].cpu(), torch.from_numpy(c["out"]))


@pytest.mark.parametrize("B,H,W", [(3, 322, 322), (2, 480, 640), (1, 337, 501), (2, 15, 29)])
def test_preprocess_sizes(u, B, H, W):
    g = torch.Generator().manual_seed(H * W)
    img = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    out = u.preprocess_images(img.cuda())
    ref = torch.stack([ao.preprocess(i) for i in img])
    assert torch.equal(out.cpu(), ref)
    # custom statistics and patch size
    out2 = u.preprocess_images(img, mean=(0.5, 0.4, 0.3), std=(0.2, 0.25, 0.5), patch=8)
    ref2 = torch.stack([ao.preprocess(i, mean=(0.5, 0.4, 0.3), std=(0.2, 0.25, 0.5), patch=8) for i in img])
    assert torch.equal(out2.cpu(), ref2)


@pytest.mark.parametrize("name", sorted(load_cases("preprocess_resize.npz")))
def test_preprocess_resize_golden(u, name):
    """ToTensor + Normalize + torchvision's antialiased tensor resize + centre crop, golden vectors made by torchvision
    (dvgl_benchmark/datasets_ws.py:233-235; demo/anyloc_vlad_generate.py:165-177)."""
    c = load_cases("preprocess_resize.npz")[name]
    mode = name.split("_")[0]
    out = u.preprocess_images(c["img"], resize=tuple(int(v) for v in c["size"]), interpolation=mode)
    assert out.shape == (1,) + c["out"].shape
    err = float((out[0].cpu() - torch.from_numpy(c["out"])).abs().max())
    assert err < 2e-5, (name, err)               # values are O(1): fp32 rounding of the weights / summation order


@pytest.mark.parametrize("mode", ["bilinear", "bicubic"])
@pytest.mark.parametrize("B,H,W,size", [(2, 720, 1280, (480, 640)), (1, 333, 517, (480, 640)), (2, 1200, 900, (1024, 768))])
def test_preprocess_resize_sizes(u, mode, B, H, W, size):
    g = torch.Generator().manual_seed(H + W)
    img = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    out = u.preprocess_images(img, resize=size, interpolation=mode)
    ref = torch.stack([ao.preprocess(i, resize=size, interpolation=mode) for i in img])
    assert out.shape == ref.shape
    assert float((out.cpu() - ref).abs().max()) < 2e-5


def test_preprocess_feeds_extractor(u):
    """uint8 images -> preprocess_images -> extractor equals the float path on the torchvision-style input."""
    from oracle import dinov2_restated as dr
    model = dr.perturb(dr.build("dinov2_vits14", seed=0, depth_override=2), seed=1)
    img = torch.randint(0, 256, (2, 60, 75, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3))
    ext = u.DinoV2ExtractFeatures("dinov2_vits14", 1, "value", device="cuda", weights=model.state_dict())
    a = ext(u.preprocess_images(img))
    b = ext(torch.stack([ao.preprocess(i) for i in img]).cuda())
    assert torch.equal(a, b)


def test_preprocess_errors(u):
    with pytest.raises(TypeError):
        u.preprocess_images(torch.zeros(1, 20, 20, 3))
    with pytest.raises(ValueError):
        u.preprocess_images(torch.zeros(1, 3, 20, 20, dtype=torch.uint8))
    with pytest.raises(ValueError):
        u.preprocess_images(torch.zeros(1, 10, 20, 3, dtype=torch.uint8))
