"""Native (hand-written sm_100a) executors, one per model family.

``builder_for(module)`` returns ``build(module, device, **kw) -> executor`` when the
wrapped module is a family we have a static kernel schedule for, else ``None`` (the
engine then falls back to a torch replica of the module).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch.nn as nn


WORKSPACE_LIMIT = 6     # distinct (batch, resolution, text length) activation workspaces kept per executor


def cache_workspace(cache: dict, key, ws: dict, device=None, on_evict: Optional[Callable] = None) -> dict:
    """Insert ``ws`` into an executor's workspace cache, evicting the least recently used entry beyond
    ``WORKSPACE_LIMIT`` (a ComfyUI session that walks through many resolutions must not pin gigabytes of activation
    buffers per shape).  ``on_evict`` lets the executor drop anything that still points into the evicted buffers
    (captured CUDA graphs); the device is synchronised first so no in-flight kernel uses them."""
    cache[key] = ws
    while len(cache) > WORKSPACE_LIMIT:
        old = next(iter(cache))
        if old == key:
            break
        if device is not None:
            import torch
            if torch.cuda.is_available():
                torch.cuda.synchronize(device)
        if on_evict is not None:
            on_evict(old)
        del cache[old]
    return ws


def touch_workspace(cache: dict, key):
    """LRU hit: move ``key`` to the most-recent end and return its workspace (or None)."""
    ws = cache.get(key)
    if ws is not None and next(reversed(cache)) != key:
        cache[key] = cache.pop(key)
    return ws


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
