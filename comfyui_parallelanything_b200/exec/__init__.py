"""Native (hand-written sm_100a) executors, one per model family.

``builder_for(module)`` returns ``build(module, device, **kw) -> executor`` when the
wrapped module is a family we have a static kernel schedule for, else ``None`` (the
engine then falls back to a torch replica of the module).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch.nn as nn


def builder_for(module: nn.Module) -> Optional[Callable]:
    fam = getattr(module, "pa_family", None)
    if fam == "flux":
        from ..models.flux import Flux
        if isinstance(module, Flux) and module.params.hidden_size // module.params.num_heads == 128 \
                and module.params.patch_size == 2:
            from .flux_exec import build_flux_executor
            return build_flux_executor
    if fam == "wan":
        from ..models.wan import WanModel
        p = getattr(module, "params", None)
        if isinstance(module, WanModel) and p.dim // p.num_heads == 128 and tuple(p.patch_size) == (1, 2, 2) \
                and p.in_dim == 16:
            from .wan_exec import build_wan_executor
            return build_wan_executor
    if fam == "zimage":
        from ..models.zimage import ZImageModel
        p = getattr(module, "params", None)
        if isinstance(module, ZImageModel) and p.dim // p.n_heads == 128 and p.patch_size == 2 and p.in_channels == 16:
            from .zimage_exec import build_zimage_executor
            return build_zimage_executor
    if fam == "vae":
        from ..models.vae import VAEDecoder
        if isinstance(module, VAEDecoder):
            from .vae_exec import build_vae_executor
            return build_vae_executor
    if fam == "unet":
        from . import unet_exec
        if unet_exec.supports(module):
            return unet_exec.build_unet_executor
    return None
