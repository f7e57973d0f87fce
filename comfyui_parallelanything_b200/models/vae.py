"""SD-style VAE decoder (latent -> RGB) in plain PyTorch, random init.

Oracle for the sm_100a VAE-decode kernels (conv3x3 implicit GEMM, GroupNorm+SiLU,
nearest-upsample) and the model used by the batch-split VAE decode path: decode is
embarrassingly parallel over the batch, so it is replicated and split exactly like
the UNet/DiT forward (SURVEY.md §2.6 "VAE decode conv").
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


class VaeResBlock(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.nin_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.nin_shortcut is not None:
            x = self.nin_shortcut(x)
        return x + h


class VaeAttnBlock(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.norm = nn.GroupNorm(32, ch, eps=1e-6)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(ch, ch, 1) for _ in range(4))

    def forward(self, x):
        b, c, h, w = x.shape
        n = self.norm(x)
        q = self.q(n).flatten(2).transpose(1, 2)[:, None]
        k = self.k(n).flatten(2).transpose(1, 2)[:, None]
        v = self.v(n).flatten(2).transpose(1, 2)[:, None]
        o = F.scaled_dot_product_attention(q, k, v)[:, 0].transpose(1, 2).reshape(b, c, h, w)
        return x + self.proj_out(o)


class VaeUpsample(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class VaeUpLevel(nn.Module):
    def __init__(self, cin: int, cout: int, n_blocks: int, upsample: bool):
        super().__init__()
        self.block = nn.ModuleList([VaeResBlock(cin if i == 0 else cout, cout) for i in range(n_blocks)])
        self.upsample = VaeUpsample(cout) if upsample else None

    def forward(self, x):
        for b in self.block:
            x = b(x)
        return self.upsample(x) if self.upsample is not None else x


class VaeMid(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.block_1 = VaeResBlock(ch, ch)
        self.attn_1 = VaeAttnBlock(ch)
        self.block_2 = VaeResBlock(ch, ch)

    def forward(self, x):
        return self.block_2(self.attn_1(self.block_1(x)))


class VAEDecoder(nn.Module):
    pa_family = "vae"

    def __init__(self, z_channels: int = 4, ch: int = 128, ch_mult: Sequence[int] = (1, 2, 4, 4),
                 num_res_blocks: int = 2, out_ch: int = 3, scaling_factor: float = 0.18215,
                 dtype: Optional[torch.dtype] = None, device=None, **_ignored):
        super().__init__()
        self.config = dict(z_channels=z_channels, ch=ch, ch_mult=list(ch_mult), num_res_blocks=num_res_blocks,
                           out_ch=out_ch, scaling_factor=scaling_factor)
        self.scaling_factor = scaling_factor
        top = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, top, 3, padding=1)
        self.mid = VaeMid(top)
        levels = []
        cin = top
        for i, m in reversed(list(enumerate(ch_mult))):
            levels.append(VaeUpLevel(cin, ch * m, num_res_blocks + 1, upsample=i != 0))
            cin = ch * m
        self.up = nn.ModuleList(levels)            # stored high->low resolution order of execution
        self.norm_out = nn.GroupNorm(32, cin, eps=1e-6)
        self.conv_out = nn.Conv2d(cin, out_ch, 3, padding=1)
        if dtype is not None or device is not None:
            self.to(device=device, dtype=dtype)

    def forward(self, z, timesteps=None, context=None, **kwargs):
        """``timesteps``/``context`` are accepted (and ignored) so the decoder can be
        driven through the same ``forward(x, timesteps, context=None, **kw)`` hook as
        the diffusion models."""
        h = self.conv_in(z / self.scaling_factor)
        h = self.mid(h)
        for lvl in self.up:
            h = lvl(h)
        return self.conv_out(F.silu(self.norm_out(h)))


def tiny_config() -> dict:
    return dict(z_channels=4, ch=32, ch_mult=[1, 2], num_res_blocks=1, out_ch=3)
