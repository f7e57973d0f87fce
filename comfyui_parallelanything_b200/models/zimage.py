"""Z-Image / Lumina-2 style single-stream DiT ("NextDiT") in plain PyTorch (random-init architecture definition).

The reference lists Z_IMAGE next to FLUX.1 and WAN2.2 as tested models (/root/reference/README.md) and its layer-split
mode walks the ``layers`` block list such models expose (/root/reference/any_device_parallel.py:1156).  As for the
other families (``models/flux.py``), ComfyUI is not installable offline, so the public architecture is defined here
with ComfyUI attribute names (``x_embedder``, ``cap_embedder``, ``t_embedder``, ``noise_refiner``,
``context_refiner``, ``layers``, ``final_layer``): the numerics oracle of ``exec/zimage_exec.py``, a model the
reference's cloning / wrapping works on, and a stock-torch module the bench's reference arm can wrap.

Architecture (Z-Image-Turbo sizes in ``zimage_turbo_params``: dim 3840, 30 heads of 128, 30 layers, SwiGLU 10240):

  * caption features [B, Lc, cap_feat_dim] -> RMSNorm + Linear -> 2 un-modulated ``context_refiner`` blocks
  * latent [B, 16, H, W] -> 2x2 patches (features ordered ph, pw, c) -> Linear -> 2 modulated ``noise_refiner`` blocks
  * ``layers``: joint blocks over the concatenation [caption | image] (one stream, shared weights)
  * every block: sandwich RMSNorms around attention (per-head q/k RMSNorm, 3-axis RoPE) and around a SwiGLU FFN;
    AdaLN: ``x + tanh(gate) * norm2(f(norm1(x) * (1 + scale)))`` with (scale, gate) x 2 from ``silu(t_emb)``
  * final layer: LayerNorm (no affine) * (1 + scale) -> Linear -> unpatchify

Forward contract: ``model(x[B,16,H,W], timesteps[B], context=cap_feats[B,Lc,cap_feat_dim])`` -> ``[B,16,H,W]``.
``timesteps`` is the DiT-native time in [0, 1] (ComfyUI's wrapper passes ``1 - sigma`` and negates the output; both
are sampler-side conventions and are not part of this module).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .flux import EmbedND, apply_rope, timestep_embedding


@dataclass
class ZImageParams:
    patch_size: int = 2
    in_channels: int = 16
    dim: int = 3840
    n_layers: int = 30
    n_refiner_layers: int = 2
    n_heads: int = 30
    ffn_hidden: int = 10240
    norm_eps: float = 1e-5
    cap_feat_dim: int = 2560
    axes_dims: List[int] = field(default_factory=lambda: [32, 48, 48])
    rope_theta: float = 256.0
    t_scale: float = 1000.0
    adaln_dim: int = 256          # min(dim, 256) in the public model


def zimage_turbo_params() -> ZImageParams:
    return ZImageParams()


def zimage_tiny_params(dim: int = 256, heads: int = 2, layers: int = 2) -> ZImageParams:
    """head_dim stays 128 (what the sm_100a kernels are specialised for)."""
    return ZImageParams(dim=dim, n_layers=layers, n_refiner_layers=1, n_heads=heads, ffn_hidden=3 * dim,
                        cap_feat_dim=192, adaln_dim=128)


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)).to(x.dtype) * self.weight


class JointAttention(nn.Module):
    def __init__(self, dim: int, n_heads: int, eps: float):
        super().__init__()
        self.n_heads, self.head_dim = n_heads, dim // n_heads
        self.qkv = nn.Linear(dim, 3 * dim, bias=False)
        self.out = nn.Linear(dim, dim, bias=False)
        self.q_norm = RMSNorm(self.head_dim, eps)
        self.k_norm = RMSNorm(self.head_dim, eps)

    def forward(self, x, pe):
        b, l, _ = x.shape
        q, k, v = self.qkv(x).view(b, l, 3, self.n_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k = apply_rope(self.q_norm(q), self.k_norm(k), pe)
        o = F.scaled_dot_product_attention(q, k, v)
        return self.out(o.transpose(1, 2).reshape(b, l, -1))


class FeedForward(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.w1 = nn.Linear(dim, hidden, bias=False)
        self.w2 = nn.Linear(hidden, dim, bias=False)
        self.w3 = nn.Linear(dim, hidden, bias=False)

    def forward(self, x):
        return self.w2(F.silu(self.w1(x)) * self.w3(x))


class JointTransformerBlock(nn.Module):
    def __init__(self, p: ZImageParams, modulation: bool = True):
        super().__init__()
        self.modulation = modulation
        self.attention = JointAttention(p.dim, p.n_heads, p.norm_eps)
        self.feed_forward = FeedForward(p.dim, p.ffn_hidden)
        self.attention_norm1 = RMSNorm(p.dim, p.norm_eps)
        self.attention_norm2 = RMSNorm(p.dim, p.norm_eps)
        self.ffn_norm1 = RMSNorm(p.dim, p.norm_eps)
        self.ffn_norm2 = RMSNorm(p.dim, p.norm_eps)
        if modulation:
            self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(p.adaln_dim, 4 * p.dim, bias=True))

    def forward(self, x, pe, adaln_input: Optional[torch.Tensor] = None):
        if self.modulation:
            scale_msa, gate_msa, scale_mlp, gate_mlp = self.adaLN_modulation(adaln_input).unsqueeze(1).chunk(4, dim=-1)
            x = x + gate_msa.tanh() * self.attention_norm2(self.attention(self.attention_norm1(x) * (1 + scale_msa), pe))
            x = x + gate_mlp.tanh() * self.ffn_norm2(self.feed_forward(self.ffn_norm1(x) * (1 + scale_mlp)))
        else:
            x = x + self.attention_norm2(self.attention(self.attention_norm1(x), pe))
            x = x + self.ffn_norm2(self.feed_forward(self.ffn_norm1(x)))
        return x


class TimestepEmbedder(nn.Module):
    def __init__(self, out_dim: int, mid_dim: int = 1024, freq_dim: int = 256):
        super().__init__()
        self.freq_dim = freq_dim
        self.mlp = nn.Sequential(nn.Linear(freq_dim, mid_dim), nn.SiLU(), nn.Linear(mid_dim, out_dim))

    def forward(self, t, t_scale: float):
        return self.mlp(timestep_embedding(t, self.freq_dim, time_factor=t_scale).to(self.mlp[0].weight.dtype))


class FinalLayer(nn.Module):
    def __init__(self, p: ZImageParams):
        super().__init__()
        self.norm_final = nn.LayerNorm(p.dim, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(p.dim, p.patch_size * p.patch_size * p.in_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(p.adaln_dim, p.dim, bias=True))

    def forward(self, x, c):
        scale = self.adaLN_modulation(c).unsqueeze(1)
        return self.linear(self.norm_final(x) * (1 + scale))


class ZImageModel(nn.Module):
    pa_family = "zimage"

    def __init__(self, params: ZImageParams, dtype=None, device=None):
        super().__init__()
        p = self.params = params
        self.dtype = dtype
        self.patch_size, self.in_channels, self.out_channels = p.patch_size, p.in_channels, p.in_channels
        with torch.device(device) if device is not None else _nullctx():
            self.x_embedder = nn.Linear(p.patch_size * p.patch_size * p.in_channels, p.dim, bias=True)
            self.noise_refiner = nn.ModuleList([JointTransformerBlock(p, True) for _ in range(p.n_refiner_layers)])
            self.context_refiner = nn.ModuleList([JointTransformerBlock(p, False) for _ in range(p.n_refiner_layers)])
            self.t_embedder = TimestepEmbedder(p.adaln_dim)
            self.cap_embedder = nn.Sequential(RMSNorm(p.cap_feat_dim, p.norm_eps), nn.Linear(p.cap_feat_dim, p.dim, bias=True))
            self.layers = nn.ModuleList([JointTransformerBlock(p, True) for _ in range(p.n_layers)])
            self.final_layer = FinalLayer(p)
        self.rope_embedder = EmbedND(p.dim // p.n_heads, p.rope_theta, p.axes_dims)
        if dtype is not None:
            self.to(dtype)
        with torch.no_grad():                    # non-trivial norm weights so the numerics checks exercise them
            for n_, p_ in self.named_parameters():
                if n_.endswith("norm1.weight") or n_.endswith("norm2.weight") or n_.endswith("_norm.weight"):
                    p_.add_(0.1 * torch.randn_like(p_))

    @staticmethod
    def make_ids(batch: int, cap_len: int, h: int, w: int, device) -> torch.Tensor:
        """[B, cap_len + h*w, 3]: caption token i sits at (i + 1, 0, 0), image token (r, c) at (cap_len + 1, r, c)."""
        cap = torch.zeros(cap_len, 3, device=device)
        cap[:, 0] = torch.arange(1, cap_len + 1, device=device)
        img = torch.zeros(h, w, 3, device=device)
        img[..., 0] = cap_len + 1
        img[..., 1] = torch.arange(h, device=device)[:, None]
        img[..., 2] = torch.arange(w, device=device)[None, :]
        return torch.cat([cap, img.reshape(h * w, 3)], 0)[None].repeat(batch, 1, 1)

    def forward(self, x, timesteps, context=None, num_tokens=None, attention_mask=None, transformer_options=None,
                **kwargs):
        p = self.params
        b, c, h, w = x.shape
        ps = p.patch_size
        hh, ww = h // ps, w // ps
        t = self.t_embedder(timesteps, p.t_scale).to(x.dtype)
        cap = self.cap_embedder(context)
        lc = cap.shape[1]
        tok = x.view(b, c, hh, ps, ww, ps).permute(0, 2, 4, 3, 5, 1).reshape(b, hh * ww, ps * ps * c)
        img = self.x_embedder(tok)
        pe = self.rope_embedder(self.make_ids(b, lc, hh, ww, x.device))
        pe_cap, pe_img = pe[:, :, :lc], pe[:, :, lc:]
        for blk in self.context_refiner:
            cap = blk(cap, pe_cap)
        for blk in self.noise_refiner:
            img = blk(img, pe_img, t)
        xs = torch.cat([cap, img], 1)
        for blk in self.layers:
            xs = blk(xs, pe, t)
        out = self.final_layer(xs[:, lc:], t)
        out = out.view(b, hh, ww, ps, ps, c).permute(0, 5, 1, 3, 2, 4).reshape(b, c, h, w)
        return out


class _nullctx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def example_inputs(params: ZImageParams, batch: int, height: int = 1024, width: int = 1024, cap_len: int = 128,
                   device="cpu", dtype=torch.bfloat16, seed: int = 0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(batch, params.in_channels, height // 8, width // 8, generator=g).to(device=device, dtype=dtype)
    t = torch.rand(batch, generator=g).to(device=device, dtype=dtype)
    ctx = torch.randn(batch, cap_len, params.cap_feat_dim, generator=g).to(device=device, dtype=dtype)
    return dict(x=x, timesteps=t, context=ctx)
