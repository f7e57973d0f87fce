"""Weight ingestion / checkpointing of *packed* executor weights (SURVEY §5 "Checkpoint / resume").

The reference has no checkpointing of its own (inference only; weights arrive inside the ComfyUI MODEL).  Our
executors re-pack weights at build time (concatenated modulation tables, fused QKV matrices, interleaved GEGLU
rows, tap-major conv weights, MXFP8 bytes + scale chunks); packing FLUX-dev takes seconds and fp8 quantisation
more, so the packed form can be saved once and memory-mapped back:

    save_packed(executor, "flux_dev.pa")            # after building from a torch module / state_dict
    load_packed_into(executor, "flux_dev.pa")       # resume: overwrite the executor's tensors in place

plus ``ingest_state_dict`` to build the plain-torch architecture from a ComfyUI-style ``state_dict`` (or a
``.safetensors`` file when the ``safetensors`` package is importable) before packing.
"""
from __future__ import annotations

import os
from typing import Any, Dict

import torch


def _collect(executor) -> Dict[str, Any]:
    w = getattr(executor, "W", None)
    if isinstance(w, dict):
        return w
    raise TypeError(f"{type(executor).__name__} does not expose a packed weight table")


def save_packed(executor, path: str) -> int:
    """Write every packed tensor (bf16 / fp8 bytes / scale chunks / ints) to ``path``; returns bytes written."""
    table = _collect(executor)
    blob = {k: (v.detach().cpu() if isinstance(v, torch.Tensor) else v) for k, v in table.items()}
    meta = {"class": type(executor).__name__, "fp8": bool(getattr(executor, "fp8", False)),
            "keys": len(blob), "format": 1}
    tmp = path + ".tmp"
    torch.save({"meta": meta, "weights": blob}, tmp)
    os.replace(tmp, path)                       # atomic: a crash never leaves a truncated checkpoint
    return os.path.getsize(path)


def load_packed_into(executor, path: str, strict: bool = True) -> Dict[str, Any]:
    """Copy a saved packed table into ``executor`` (same architecture / fp8 mode).  Returns the meta dict."""
    ck = torch.load(path, map_location="cpu", mmap=True, weights_only=False)
    meta, blob = ck["meta"], ck["weights"]
    table = _collect(executor)
    if strict:
        if meta.get("class") != type(executor).__name__:
            raise ValueError(f"checkpoint is for {meta.get('class')}, not {type(executor).__name__}")
        missing = [k for k in table if k not in blob]
        extra = [k for k in blob if k not in table]
        if missing or extra:
            raise KeyError(f"packed table mismatch: missing {missing[:3]}, unexpected {extra[:3]}")
    with torch.no_grad():
        for k, v in blob.items():
            cur = table.get(k)
            if isinstance(cur, torch.Tensor) and isinstance(v, torch.Tensor):
                if cur.shape != v.shape or cur.dtype != v.dtype:
                    raise ValueError(f"{k}: shape/dtype mismatch {tuple(cur.shape)}/{cur.dtype} vs "
                                     f"{tuple(v.shape)}/{v.dtype}")
                cur.copy_(v, non_blocking=True)
            elif not strict or k in table:
                table[k] = v
    return meta


def ingest_state_dict(model: torch.nn.Module, source, prefix: str = "", strict: bool = False):
    """Load a ``state_dict`` (dict, ``.pt``/``.pth`` path or ``.safetensors`` path) into the plain-torch
    architecture.  ``prefix`` strips e.g. ``model.diffusion_model.`` from ComfyUI checkpoints."""
    if isinstance(source, (str, os.PathLike)):
        p = str(source)
        if p.endswith(".safetensors"):
            try:
                from safetensors.torch import load_file  # type: ignore
            except ImportError as e:  # pragma: no cover - optional dependency
                raise RuntimeError("reading .safetensors needs the `safetensors` package") from e
            sd = load_file(p)
        else:
            sd = torch.load(p, map_location="cpu", weights_only=True)
    else:
        sd = source
    if prefix:
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    return model.load_state_dict(sd, strict=strict)
