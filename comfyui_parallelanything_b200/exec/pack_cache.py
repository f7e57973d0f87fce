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


def packed_table(executor) -> Dict[str, Any]:
    """name -> packed tensor (or small int/None) of an executor.  DiT executors keep a flat ``W`` dict; the UNet / VAE
    executors hold their packed weights in small per-layer objects (``_Conv`` / ``_Lin`` / ``_GN`` ...), which are walked
    into dotted paths (``inp.3.0.1.conv1.w``) - deterministic for a given architecture, so a table saved from one
    executor loads into another built from the same model definition."""
    w = getattr(executor, "W", None)
    if isinstance(w, dict):
        return w
    table: Dict[str, Any] = {}
    seen = set()

    def walk(o, path):
        if isinstance(o, torch.Tensor):
            if id(o) not in seen:
                seen.add(id(o))
                table[path] = o
        elif isinstance(o, (list, tuple)):
            for i, v in enumerate(o):
                walk(v, f"{path}.{i}")
        elif isinstance(o, dict):
            for k in sorted(o, key=str):
                walk(o[k], f"{path}.{k}")
        elif hasattr(o, "__dict__") and not isinstance(o, (torch.nn.Module, type)):
            for k in sorted(vars(o)):
                if not k.startswith("_"):
                    walk(getattr(o, k), f"{path}.{k}")
    for name in sorted(vars(executor)):
        if name.startswith("_") or name in ("device", "training"):
            continue
        v = getattr(executor, name)
        if isinstance(v, (torch.Tensor, list, tuple, dict)) or (hasattr(v, "__dict__")
                                                                and not isinstance(v, (torch.nn.Module, type))):
            walk(v, name)
    if not table:
        raise TypeError(f"{type(executor).__name__} does not expose any packed weights")
    return table


_collect = packed_table

FORMAT = 2


def save_packed(executor, path: str) -> int:
    """Write every packed tensor (bf16 / fp8 bytes / scale chunks / ints) to ``path``; returns bytes written."""
    table = packed_table(executor)
    blob = {k: (v.detach().cpu() if isinstance(v, torch.Tensor) else v) for k, v in table.items()}
    meta = {"class": type(executor).__name__, "fp8": bool(getattr(executor, "fp8", False)),
            "keys": len(blob), "format": FORMAT}
    tmp = path + ".tmp"
    torch.save({"meta": meta, "weights": blob}, tmp)
    os.replace(tmp, path)                       # atomic: a crash never leaves a truncated checkpoint
    return os.path.getsize(path)


def load_packed_into(executor, path: str, strict: bool = True) -> Dict[str, Any]:
    """Copy a saved packed table into ``executor`` (same architecture / fp8 mode).  Returns the meta dict.
    The file is read with ``weights_only=True``: a ``.pa`` checkpoint holds tensors, ints, None and a small meta
    dict only, so unpickling never executes code from the file."""
    ck = torch.load(path, map_location="cpu", mmap=True, weights_only=True)
    meta, blob = ck["meta"], ck["weights"]
    if meta.get("format") not in (1, FORMAT):
        raise ValueError(f"unknown packed-checkpoint format {meta.get('format')!r}")
    table = packed_table(executor)
    flat = isinstance(getattr(executor, "W", None), dict)
    if strict:
        if meta.get("class") != type(executor).__name__:
            raise ValueError(f"checkpoint is for {meta.get('class')}, not {type(executor).__name__}")
        if bool(meta.get("fp8", False)) != bool(getattr(executor, "fp8", False)):
            raise ValueError(f"checkpoint fp8={meta.get('fp8')} but the executor was built with "
                             f"fp8={bool(getattr(executor, 'fp8', False))}")
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
            elif flat and (not strict or k in table):
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
