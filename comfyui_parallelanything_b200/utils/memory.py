"""Memory hygiene + model-attribute scrubbing helpers.

Behavioural parity with /root/reference/any_device_parallel.py:
  * ``disable_flash_xformers``  ADP:126-164
  * ``clear_model_caches``      ADP:166-195 (``clear_flux_caches``)
  * ``aggressive_cleanup``      ADP:197-209
  * ``get_free_vram``           ADP:724-735 — fixed: uses ``cudaMemGetInfo`` (true free
    memory) and does not change the caller's current device.
The ComfyUI hooks (``soft_empty_cache`` / ``unload_all_models``) are optional:
ComfyUI is not importable in this image, so they are looked up lazily.
"""
from __future__ import annotations

import gc
from typing import Iterable, Optional

import torch
import torch.nn as nn

from . import log

CACHE_ATTRS = (
    "img_ids", "txt_ids", "_img_ids", "_txt_ids", "cached_img_ids", "cached_txt_ids",
    "pos_emb", "_pos_emb", "pos_embed", "_pos_embed", "cached_pos_emb",
    "rope", "_rope", "freqs_cis", "_freqs_cis", "freqs", "_freqs",
    "cache", "_cache", "kv_cache", "_kv_cache", "attn_bias", "_attn_bias",
    "rope_cache", "_rope_cache", "freqs_cis_cache", "_freqs_cis_cache",
    "temporal_ids", "frame_ids", "video_ids", "temp_pos_emb",
)

_ATTN_TOGGLES_MODEL = (
    ("set_use_memory_efficient_attention_xformers", False),
    ("set_use_flash_attention_2", False),
    ("disable_xformers_memory_efficient_attention", None),
    ("use_xformers", False),
    ("use_flash_attention", False),
    ("use_flash_attention_2", False),
    ("_use_memory_efficient_attention", False),
    ("_flash_attention_enabled", False),
)
_ATTN_FLAGS_SUBMODULE = (
    "use_xformers", "use_flash_attention", "use_flash_attention_2",
    "_use_memory_efficient_attention", "enable_flash", "enable_xformers",
)


def comfy_mm():
    """Return ``comfy.model_management`` if ComfyUI is importable, else None."""
    try:
        import comfy.model_management as mm  # type: ignore
        return mm
    except Exception:
        return None


def disable_flash_xformers(model: nn.Module) -> int:
    """Best-effort switch to plain attention on ``model`` (for < sm_80 replicas).
    Returns the number of toggles applied.  Never raises."""
    n = 0
    for name, value in _ATTN_TOGGLES_MODEL:
        if not hasattr(model, name):
            continue
        try:
            target = getattr(model, name)
            if callable(target):
                target() if value is None else target(value)
            else:
                setattr(model, name, value)
            n += 1
        except Exception:
            pass
    for mod_name, module in model.named_modules():
        low = mod_name.lower()
        if not any(tag in low for tag in ("attn", "attention", "transformer")):
            continue
        for flag in _ATTN_FLAGS_SUBMODULE:
            if hasattr(module, flag):
                try:
                    setattr(module, flag, False)
                    n += 1
                except AttributeError:
                    pass
        if hasattr(module, "set_processor"):
            try:
                from diffusers.models.attention_processor import Attention  # type: ignore
                module.set_processor(Attention())
                n += 1
            except Exception:
                pass
    return n


def clear_model_caches(model: nn.Module, attrs: Iterable[str] = CACHE_ATTRS, quiet: bool = False) -> int:
    """Null device-bound cached tensors (rope tables, ids, kv caches, ...) on the
    model and every submodule so they don't follow a clone to another GPU."""
    cleared = 0
    seen = set()
    objs = [model] + [m for _, m in model.named_modules()]
    for obj in objs:
        if id(obj) in seen:
            continue
        seen.add(id(obj))
        for a in attrs:
            # never touch registered submodules/params/buffers that happen to share a name
            if a in getattr(obj, "_modules", {}) or a in getattr(obj, "_parameters", {}):
                continue
            if a in getattr(obj, "_buffers", {}):
                continue
            try:
                if getattr(obj, a, None) is not None:
                    setattr(obj, a, None)
                    cleared += 1
            except (AttributeError, TypeError):
                pass
    if cleared and not quiet:
        log.info("Cleared %d cached tensors", cleared)
    return cleared


# reference-compatible alias (ADP:166)
clear_flux_caches = clear_model_caches


def aggressive_cleanup() -> None:
    gc.collect()
    if torch.cuda.is_available():
        for i in range(torch.cuda.device_count()):
            try:
                with torch.cuda.device(i):
                    torch.cuda.synchronize()
                    torch.cuda.empty_cache()
            except Exception:
                pass
    mm = comfy_mm()
    if mm is not None:
        try:
            mm.soft_empty_cache()
        except Exception:
            pass


def get_free_vram(device_name: str) -> float:
    """Free device memory in MiB (0 for non-CUDA / unknown devices)."""
    try:
        d = torch.device(device_name)
        if d.type != "cuda" or not torch.cuda.is_available():
            return 0.0
        idx = d.index if d.index is not None else torch.cuda.current_device()
        free, _total = torch.cuda.mem_get_info(idx)
        return free / float(1024 ** 2)
    except Exception:
        return 0.0


def total_vram(device_name: str) -> float:
    try:
        d = torch.device(device_name)
        if d.type != "cuda" or not torch.cuda.is_available():
            return 0.0
        return torch.cuda.get_device_properties(d).total_memory / float(1024 ** 2)
    except Exception:
        return 0.0


def module_device(module: nn.Module) -> Optional[torch.device]:
    for p in module.parameters():
        return p.device
    for b in module.buffers():
        return b.device
    return None


def module_bytes(module: nn.Module) -> int:
    seen, total = set(), 0
    for t in list(module.parameters()) + list(module.buffers()):
        if t.data_ptr() in seen:
            continue
        seen.add(t.data_ptr())
        total += t.numel() * t.element_size()
    return total
