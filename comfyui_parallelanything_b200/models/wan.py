"""WAN2.x-style video DiT in plain PyTorch, random init (stock torch ops only).

Shape of the public WAN 2.1/2.2 T2V transformer as ComfyUI instantiates it:
``patch_embedding`` Conv3d(1,2,2), text/time embedders, ``blocks`` of
(self-attn with full-width RMS q/k norm + 3-D RoPE, cross-attn to the text,
GELU-tanh FFN) modulated by a 6-way time projection, and a modulated head.
WAN2.2-A14B uses dim 5120 / ffn 13824 / 40 heads / 40 layers per expert.

Note the block list is called ``blocks`` — the reference's pipeline planner only
looks at ``double_blocks/single_blocks/transformer_blocks/layers``
(/root/reference/any_device_parallel.py:1156), so its batch==1 mode is a no-op for
WAN; ours includes ``blocks`` (parallel/pipeline.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .flux import EmbedND, apply_rope, timestep_embedding


@dataclass
class WanParams:
    in_dim: int = 16
    out_dim: int = 16
    dim: int = 5120
    ffn_dim: int = 13824
    num_heads: int = 40
    num_layers: int = 40
    text_dim: int = 4096
    text_len: int = 512
    freq_dim: int = 256
    patch_size: Tuple[int, int, int] = (1, 2, 2)
    eps: float = 1e-6


def wan22_a14b_params() -> WanParams:
    return WanParams()


def wan_tiny_params() -> WanParams:
    return WanParams(in_dim=16, out_dim=16, dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64,
                     text_len=16)


class WanRMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)).type_as(x) * self.weight


class WanSelfAttention(nn.Module):
    def __init__(self, dim: int, num_heads: int, eps: float):
        super().__init__()
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.q, self.k, self.v, self.o = (nn.Linear(dim, dim) for _ in range(4))
        self.norm_q, self.norm_k = WanRMSNorm(dim, eps), WanRMSNorm(dim, eps)

    def forward(self, x, freqs):
        b, s, _ = x.shape
        n, d = self.num_heads, self.head_dim
        q = self.norm_q(self.q(x)).view(b, s, n, d).transpose(1, 2)
        k = self.norm_k(self.k(x)).view(b, s, n, d).transpose(1, 2)
        v = self.v(x).view(b, s, n, d).transpose(1, 2)
        q, k = apply_rope(q, k, freqs)
        o = F.scaled_dot_product_attention(q, k, v)
        return self.o(o.transpose(1, 2).reshape(b, s, n * d))


class WanCrossAttention(WanSelfAttention):
    def forward(self, x, context):  # type: ignore[override]
        b, s, _ = x.shape
        n, d = self.num_heads, self.head_dim
        q = self.norm_q(self.q(x)).view(b, s, n, d).transpose(1, 2)
        k = self.norm_k(self.k(context)).view(b, -1, n, d).transpose(1, 2)
        v = self.v(context).view(b, -1, n, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        return self.o(o.transpose(1, 2).reshape(b, s, n * d))


class WanAttentionBlock(nn.Module):
    def __init__(self, dim: int, ffn_dim: int, num_heads: int, eps: float):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps, elementwise_affine=False)
        self.self_attn = WanSelfAttention(dim, num_heads, eps)
        self.norm3 = nn.LayerNorm(dim, eps, elementwise_affine=True)
        self.cross_attn = WanCrossAttention(dim, num_heads, eps)
        self.norm2 = nn.LayerNorm(dim, eps, elementwise_affine=False)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn_dim), nn.GELU(approximate="tanh"), nn.Linear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)

    def forward(self, x, e, freqs, context):
        e = (self.modulation.to(e.dtype) + e).chunk(6, dim=1)
        y = self.self_attn(self.norm1(x) * (1 + e[1]) + e[0], freqs)
        x = x + y * e[2]
        x = x + self.cross_attn(self.norm3(x), context)
        y = self.ffn(self.norm2(x) * (1 + e[4]) + e[3])
        return x + y * e[5]


class Head(nn.Module):
    def __init__(self, dim: int, out_dim: int, patch_size, eps: float):
        super().__init__()
        self.norm = nn.LayerNorm(dim, eps, elementwise_affine=False)
        self.head = nn.Linear(dim, out_dim * patch_size[0] * patch_size[1] * patch_size[2])
        self.modulation = nn.Parameter(torch.randn(1, 2, dim) / dim ** 0.5)

    def forward(self, x, e):
        e = (self.modulation.to(e.dtype) + e.unsqueeze(1)).chunk(2, dim=1)
        return self.head(self.norm(x) * (1 + e[1]) + e[0])


class WanModel(nn.Module):
    pa_family = "wan"

    def __init__(self, params: Optional[WanParams] = None, dtype=None, device=None, **kw):
        super().__init__()
        if params is None:
            params = WanParams(**{k: v for k, v in kw.items() if k in WanParams.__dataclass_fields__})
        self.params = p = params
        self.dim, self.num_heads, self.patch_size = p.dim, p.num_heads, tuple(p.patch_size)
        self.freq_dim = p.freq_dim
        self.dtype = dtype
        self.patch_embedding = nn.Conv3d(p.in_dim, p.dim, kernel_size=self.patch_size, stride=self.patch_size)
        self.text_embedding = nn.Sequential(nn.Linear(p.text_dim, p.dim), nn.GELU(approximate="tanh"),
                                            nn.Linear(p.dim, p.dim))
        self.time_embedding = nn.Sequential(nn.Linear(p.freq_dim, p.dim), nn.SiLU(), nn.Linear(p.dim, p.dim))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(p.dim, p.dim * 6))
        self.blocks = nn.ModuleList([WanAttentionBlock(p.dim, p.ffn_dim, p.num_heads, p.eps)
                                     for _ in range(p.num_layers)])
        self.head = Head(p.dim, p.out_dim, self.patch_size, p.eps)
        d = p.dim // p.num_heads
        self.rope_embedder = EmbedND(dim=d, theta=10000, axes_dim=[d - 4 * (d // 6), 2 * (d // 6), 2 * (d // 6)])
        if dtype is not None or device is not None:
            self.to(device=device, dtype=dtype)

    def make_ids(self, b: int, t: int, h: int, w: int, device) -> torch.Tensor:
        ids = torch.zeros(t, h, w, 3, device=device, dtype=torch.float32)
        ids[..., 0] = torch.arange(t, device=device, dtype=torch.float32)[:, None, None]
        ids[..., 1] = torch.arange(h, device=device, dtype=torch.float32)[None, :, None]
        ids[..., 2] = torch.arange(w, device=device, dtype=torch.float32)[None, None, :]
        return ids.reshape(1, t * h * w, 3).expand(b, -1, -1)

    def unpatchify(self, x: torch.Tensor, grid) -> torch.Tensor:
        b = x.shape[0]
        c = self.params.out_dim
        pt, ph, pw = self.patch_size
        t, h, w = grid
        u = x.view(b, t, h, w, pt, ph, pw, c).permute(0, 7, 1, 4, 2, 5, 3, 6)
        return u.reshape(b, c, t * pt, h * ph, w * pw)

    def forward(self, x, timesteps, context=None, clip_fea=None, transformer_options=None, **kwargs):
        # x: [B, C, T, H, W]
        x = self.patch_embedding(x)
        grid = x.shape[2:]
        x = x.flatten(2).transpose(1, 2)
        e = self.time_embedding(timestep_embedding(timesteps, self.freq_dim, time_factor=1.0).to(x.dtype))
        e0 = self.time_projection(e).unflatten(1, (6, self.dim))
        ctx = self.text_embedding(context)
        freqs = self.rope_embedder(self.make_ids(x.shape[0], *grid, x.device))
        for blk in self.blocks:
            x = blk(x, e0, freqs, ctx)
        x = self.head(x, e)
        return self.unpatchify(x, grid)


def example_inputs(params: WanParams, batch: int, frames: int = 16, height: int = 720, width: int = 1280,
                   device="cpu", dtype=torch.bfloat16, seed: int = 0):
    """720p x 16 frames -> latent [16, 4, 90, 160] (VAE 4x temporal, 8x spatial), 14 400 tokens."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    lt = max(1, frames // 4)
    x = torch.randn(batch, params.in_dim, lt, height // 8, width // 8, generator=g).to(device=device, dtype=dtype)
    t = (torch.rand(batch, generator=g) * 1000).to(device=device, dtype=dtype)
    ctx = torch.randn(batch, params.text_len, params.text_dim, generator=g).to(device=device, dtype=dtype)
    return dict(x=x, timesteps=t, context=ctx)
