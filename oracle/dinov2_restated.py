"""TEST INFRASTRUCTURE -- CPU restatement of facebookresearch/dinov2 @ main
`DinoVisionTransformer` (eval, no register tokens, block_chunks=0), which the
reference loads through torch.hub (/root/reference/utilities.py:239-240) and
hooks at `blocks[layer]` / `blocks[layer].attn.qkv` (utilities.py:245-252).
The package is NOT vendored by the reference and cannot be fetched here, so
this file restates its published forward pass:

  prepare_tokens: Conv2d(3,D,14,14) -> flatten -> cat cls -> + pos_embed
     (bicubic resize of the 37x37 table, scale_factor=(g+0.1)/37, identity
     when the grid is 37x37 and the image square)
  block: x += g1 * proj(softmax(q k^T / sqrt(64)) v);  x += g2 * ffn(LN2(x))
  ffn  : fc2(GELU_erf(fc1(x)))              (vits14 / vitb14 / vitl14)
         w3(silu(x1) * x2), [x1,x2]=w12(x)   (vitg14, hidden 4096)
  LayerNorm eps 1e-6, qkv/proj/ffn bias, LayerScale init 1.0.

Module/parameter names follow the upstream state_dict so real hub weights load
unchanged.  Structure is cross-checked in tests against the independent
HuggingFace port (transformers.models.dinov2).  PARITY UNPINNED with respect to
reference-held vectors (the reference has none).
"""
import math

import torch
from torch import nn
from torch.nn import functional as F

ARCHS = {
    # name: (embed_dim, depth, heads, ffn)
    "dinov2_vits14": (384, 12, 6, "mlp"),
    "dinov2_vitb14": (768, 12, 12, "mlp"),
    "dinov2_vitl14": (1024, 24, 16, "mlp"),
    "dinov2_vitg14": (1536, 40, 24, "swiglufused"),
}


class PatchEmbed(nn.Module):
    def __init__(self, dim, patch=14):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.scale = (dim // heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * self.scale, qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class SwiGLUFFNFused(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        hidden = (int(hidden * 2 / 3) + 7) // 8 * 8
        self.w12 = nn.Linear(dim, 2 * hidden)
        self.w3 = nn.Linear(hidden, dim)

    def forward(self, x):
        x1, x2 = self.w12(x).chunk(2, dim=-1)
        return self.w3(F.silu(x1) * x2)


class LayerScale(nn.Module):
    def __init__(self, dim, init_values=1.0):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class Block(nn.Module):
    def __init__(self, dim, heads, ffn):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, heads)
        self.ls1 = LayerScale(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, 4 * dim) if ffn == "mlp" else SwiGLUFFNFused(dim, 4 * dim)
        self.ls2 = LayerScale(dim)

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x)))
        x = x + self.ls2(self.mlp(self.norm2(x)))
        return x


class DinoVisionTransformer(nn.Module):
    def __init__(self, name="dinov2_vits14", depth_override=None):
        super().__init__()
        dim, depth, heads, ffn = ARCHS[name]
        if depth_override is not None:
            depth = depth_override
        self.embed_dim, self.num_heads, self.patch_size = dim, heads, 14
        self.interpolate_offset = 0.1
        self.patch_embed = PatchEmbed(dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, 1 + 37 * 37, dim))
        self.mask_token = nn.Parameter(torch.zeros(1, dim))     # present upstream, unused in eval
        self.blocks = nn.ModuleList([Block(dim, heads, ffn) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.head = nn.Identity()
        self.init_weights()

    def init_weights(self):
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def interpolate_pos_encoding(self, x, w, h):
        npatch = x.shape[1] - 1
        N = self.pos_embed.shape[1] - 1
        if npatch == N and w == h:
            return self.pos_embed
        pos_embed = self.pos_embed.float()
        class_pos_embed = pos_embed[:, 0]
        patch_pos_embed = pos_embed[:, 1:]
        dim = x.shape[-1]
        w0 = w // self.patch_size
        h0 = h // self.patch_size
        M = int(math.sqrt(N))
        sx = float(w0 + self.interpolate_offset) / M
        sy = float(h0 + self.interpolate_offset) / M
        patch_pos_embed = F.interpolate(
            patch_pos_embed.reshape(1, M, M, dim).permute(0, 3, 1, 2),
            scale_factor=(sx, sy), mode="bicubic", antialias=False)
        assert (w0, h0) == tuple(patch_pos_embed.shape[-2:])
        patch_pos_embed = patch_pos_embed.permute(0, 2, 3, 1).view(1, -1, dim)
        return torch.cat((class_pos_embed.unsqueeze(0), patch_pos_embed), dim=1).to(x.dtype)

    def prepare_tokens(self, x):
        B, nc, w, h = x.shape
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)
        return x + self.interpolate_pos_encoding(x, w, h)

    def forward(self, x):
        x = self.prepare_tokens(x)
        for blk in self.blocks:
            x = blk(x)
        return self.head(self.norm(x)[:, 0])


def build(name, seed=0, dtype=torch.float32, depth_override=None):
    """Random-init model with the upstream init recipe, deterministic in `seed`
    (CPU generator), eval mode."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = DinoVisionTransformer(name, depth_override=depth_override).eval().to(dtype)
    torch.random.set_rng_state(g)
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def perturb(model, seed=1, scale=0.05):
    """Make biases / LayerNorm / LayerScale non-trivial so tests exercise them
    (upstream init leaves them at 0/1)."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("bias") or "norm" in n or n.endswith("gamma"):
                p.add_(scale * torch.randn(p.shape, generator=gen))
        model.cls_token.add_(0.02 * torch.randn(model.cls_token.shape, generator=gen))
    return model
