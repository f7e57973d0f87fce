"""Latent-diffusion UNet (SD1.5 / SDXL-base shaped) in plain PyTorch, random init.

Architecture follows the public ``openaimodel.UNetModel`` used by ComfyUI for
SD1.x/SDXL (attribute names ``input_blocks / middle_block / output_blocks``,
``time_embed``, ``label_emb``, ``unet_config``) so the reference's config
harvesting (/root/reference/any_device_parallel.py:284-350: ``model_channels``,
``channel_mult``, ``transformer_depth``, ``context_dim``, ``adm_in_channels`` ...)
sees what it expects.  Stock torch ops only; used as oracle for the sm_100a
GroupNorm+SiLU / conv / cross-attention kernels and by the bench reference arm.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


def sinusoidal_embedding(t: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class GroupNorm32(nn.GroupNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype) if x.dtype == torch.float16 else super().forward(x)


class Upsample(nn.Module):
    def __init__(self, ch: int, out_ch: Optional[int] = None):
        super().__init__()
        self.conv = nn.Conv2d(ch, out_ch or ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class Downsample(nn.Module):
    def __init__(self, ch: int, out_ch: Optional[int] = None):
        super().__init__()
        self.op = nn.Conv2d(ch, out_ch or ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.op(x)


class ResBlock(nn.Module):
    def __init__(self, ch: int, emb_ch: int, out_ch: int):
        super().__init__()
        self.in_layers = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(), nn.Conv2d(ch, out_ch, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_ch, out_ch))
        self.out_layers = nn.Sequential(GroupNorm32(32, out_ch), nn.SiLU(), nn.Dropout(0.0),
                                        nn.Conv2d(out_ch, out_ch, 3, padding=1))
        self.skip_connection = nn.Identity() if ch == out_ch else nn.Conv2d(ch, out_ch, 1)

    def forward(self, x, emb):
        h = self.in_layers(x)
        h = h + self.emb_layers(emb).type(h.dtype)[:, :, None, None]
        h = self.out_layers(h)
        return self.skip_connection(x) + h


class CrossAttention(nn.Module):
    def __init__(self, query_dim: int, context_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim or query_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(0.0))

    def forward(self, x, context=None):
        ctx = x if context is None else context
        b, n, _ = x.shape
        q = self.to_q(x).view(b, n, self.heads, self.dim_head).transpose(1, 2)
        k = self.to_k(ctx).view(b, ctx.shape[1], self.heads, self.dim_head).transpose(1, 2)
        v = self.to_v(ctx).view(b, ctx.shape[1], self.heads, self.dim_head).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        return self.to_out(o.transpose(1, 2).reshape(b, n, self.heads * self.dim_head))


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.Sequential(GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim))

    def forward(self, x):
        return self.net(x)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, context_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = CrossAttention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = CrossAttention(dim, context_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), context)
        return x + self.ff(self.norm3(x))


class SpatialTransformer(nn.Module):
    def __init__(self, ch: int, heads: int, dim_head: int, depth: int, context_dim: int, use_linear: bool):
        super().__init__()
        inner = heads * dim_head
        self.use_linear = use_linear
        self.norm = nn.GroupNorm(32, ch, eps=1e-6)
        self.proj_in = nn.Linear(ch, inner) if use_linear else nn.Conv2d(ch, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, context_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(inner, ch) if use_linear else nn.Conv2d(inner, ch, 1)

    def forward(self, x, context):
        b, c, h, w = x.shape
        x_in = x
        x = self.norm(x)
        if not self.use_linear:
            x = self.proj_in(x)
        x = x.flatten(2).transpose(1, 2)
        if self.use_linear:
            x = self.proj_in(x)
        for blk in self.transformer_blocks:
            x = blk(x, context)
        if self.use_linear:
            x = self.proj_out(x)
        x = x.transpose(1, 2).reshape(b, -1, h, w)
        if not self.use_linear:
            x = self.proj_out(x)
        return x + x_in


class TimestepEmbedSequential(nn.Sequential):
    def forward(self, x, emb, context=None):
        for layer in self:
            if isinstance(layer, ResBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            else:
                x = layer(x)
        return x


class UNetModel(nn.Module):
    pa_family = "unet"

    def __init__(self, in_channels: int = 4, model_channels: int = 320, out_channels: int = 4,
                 num_res_blocks: int = 2, channel_mult: Sequence[int] = (1, 2, 4, 4),
                 transformer_depth: Sequence[int] = (1, 1, 1, 0), context_dim: int = 768,
                 num_heads: int = 8, num_head_channels: int = -1, adm_in_channels: Optional[int] = None,
                 use_linear_in_transformer: bool = False, transformer_depth_middle: Optional[int] = None,
                 dtype: Optional[torch.dtype] = None, device=None, **_ignored):
        super().__init__()
        self.unet_config = dict(in_channels=in_channels, model_channels=model_channels, out_channels=out_channels,
                                num_res_blocks=num_res_blocks, channel_mult=list(channel_mult),
                                transformer_depth=list(transformer_depth), context_dim=context_dim,
                                num_heads=num_heads, num_head_channels=num_head_channels,
                                adm_in_channels=adm_in_channels,
                                use_linear_in_transformer=use_linear_in_transformer,
                                transformer_depth_middle=transformer_depth_middle)
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.channel_mult = num_res_blocks, list(channel_mult)
        self.transformer_depth, self.context_dim = list(transformer_depth), context_dim
        self.num_heads, self.num_head_channels = num_heads, num_head_channels
        self.adm_in_channels = adm_in_channels
        self.dtype = dtype
        mc = model_channels
        emb = mc * 4
        self.time_embed = nn.Sequential(nn.Linear(mc, emb), nn.SiLU(), nn.Linear(emb, emb))
        if adm_in_channels is not None:
            self.label_emb = nn.Sequential(nn.Sequential(nn.Linear(adm_in_channels, emb), nn.SiLU(),
                                                         nn.Linear(emb, emb)))

        def heads_for(ch: int):
            if num_head_channels > 0:
                return ch // num_head_channels, num_head_channels
            return num_heads, ch // num_heads

        def st(ch: int, depth: int):
            h, d = heads_for(ch)
            return SpatialTransformer(ch, h, d, depth, context_dim, use_linear_in_transformer)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, mc, 3, padding=1))])
        chans = [mc]
        ch = mc
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers: List[nn.Module] = [ResBlock(ch, emb, mult * mc)]
                ch = mult * mc
                if self.transformer_depth[level] > 0:
                    layers.append(st(ch, self.transformer_depth[level]))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch)))
                chans.append(ch)
        # SD1.5 has (depth-1) attention in the middle even though its last level has none
        mid_depth = transformer_depth_middle if transformer_depth_middle is not None \
            else max(1, self.transformer_depth[-1])
        mid: List[nn.Module] = [ResBlock(ch, emb, ch)]
        if mid_depth > 0:
            mid += [st(ch, mid_depth)]
        mid += [ResBlock(ch, emb, ch)]
        self.middle_block = TimestepEmbedSequential(*mid)
        self.output_blocks = nn.ModuleList()
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [ResBlock(ch + chans.pop(), emb, mult * mc)]
                ch = mult * mc
                if self.transformer_depth[level] > 0:
                    layers.append(st(ch, self.transformer_depth[level]))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch))
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(), nn.Conv2d(mc, out_channels, 3, padding=1))
        if dtype is not None or device is not None:
            self.to(device=device, dtype=dtype)

    def forward(self, x, timesteps=None, context=None, y=None, control=None, transformer_options=None, **kwargs):
        emb = self.time_embed(sinusoidal_embedding(timesteps, self.model_channels).to(x.dtype))
        if self.adm_in_channels is not None:
            if y is None:
                raise ValueError("class-conditional UNet needs y")
            emb = emb + self.label_emb(y)
        hs = []
        h = x
        for module in self.input_blocks:
            h = module(h, emb, context)
            hs.append(h)
        h = self.middle_block(h, emb, context)
        for module in self.output_blocks:
            h = torch.cat([h, hs.pop()], dim=1)
            h = module(h, emb, context)
        return self.out(h.type(x.dtype))


def sd15_config() -> dict:
    return dict(in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, channel_mult=[1, 2, 4, 4],
                transformer_depth=[1, 1, 1, 0], context_dim=768, num_heads=8, num_head_channels=-1,
                adm_in_channels=None, use_linear_in_transformer=False)


def sdxl_config() -> dict:
    return dict(in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, channel_mult=[1, 2, 4],
                transformer_depth=[0, 2, 10], context_dim=2048, num_heads=-1, num_head_channels=64,
                adm_in_channels=2816, use_linear_in_transformer=True, transformer_depth_middle=10)


def tiny_config(adm: Optional[int] = None) -> dict:
    return dict(in_channels=4, model_channels=32, out_channels=4, num_res_blocks=1, channel_mult=[1, 2],
                transformer_depth=[1, 1], context_dim=64, num_heads=2, num_head_channels=-1,
                adm_in_channels=adm, use_linear_in_transformer=adm is not None)


def mini_sdxl_config() -> dict:
    """SDXL-shaped (head_dim 64, linear transformer projections, adm conditioning) but small."""
    return dict(in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1, channel_mult=[1, 2, 2],
                transformer_depth=[0, 1, 2], context_dim=128, num_heads=-1, num_head_channels=64,
                adm_in_channels=64, use_linear_in_transformer=True, transformer_depth_middle=2)


def example_inputs(cfg: dict, batch: int, height: int, width: int, ctx_len: int = 77, device="cpu",
                   dtype=torch.float32, seed: int = 0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(batch, cfg["in_channels"], height // 8, width // 8, generator=g).to(device=device, dtype=dtype)
    t = (torch.rand(batch, generator=g) * 999).to(device=device, dtype=dtype)
    ctx = torch.randn(batch, ctx_len, cfg["context_dim"], generator=g).to(device=device, dtype=dtype)
    out = dict(x=x, timesteps=t, context=ctx)
    if cfg.get("adm_in_channels"):
        out["y"] = torch.randn(batch, cfg["adm_in_channels"], generator=g).to(device=device, dtype=dtype)
    return out
