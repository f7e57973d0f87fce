"""Batch scatter/gather helpers of the host path.

Reference behaviour: /root/reference/any_device_parallel.py:1210-1285.  Differences
that are deliberate fixes:
  * a list of tensors whose members do not all have ``shape[0]==B`` is *replicated*
    to every worker instead of being silently dropped (ADP:1259-1263);
  * dicts (``transformer_options``, ``control`` residuals) are recursed into: batch-
    sized tensors inside are split, everything is moved to the worker's device;
  * ``strict_compat=True`` reproduces the reference's behaviour bit-for-bit for the
    differential tests.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import torch

from ..utils import dtypes


def get_batch_size(x: Any) -> int:
    if isinstance(x, torch.Tensor):
        return int(x.shape[0]) if x.dim() > 0 else 1
    if isinstance(x, (list, tuple)) and len(x) > 0:
        for t in x:
            if isinstance(t, torch.Tensor):
                if isinstance(x[0], torch.Tensor):
                    return int(t.shape[0])
                break
        return len(x)
    return 1


def split_value(x: Any, sizes: Sequence[int]) -> List[Any]:
    """Split along dim 0; non-tensors are replicated (ADP:1222-1237)."""
    n = len(sizes)
    if isinstance(x, torch.Tensor):
        return list(torch.split(x, list(sizes), dim=0))
    if isinstance(x, (list, tuple)):
        cols = [split_value(t, sizes) if isinstance(t, torch.Tensor) else [t] * n for t in x]
        return [type(x)(c[i] for c in cols) for i in range(n)]
    return [x] * n


def _is_batched(t: Any, batch: int) -> bool:
    return isinstance(t, torch.Tensor) and t.dim() > 0 and t.shape[0] == batch


def split_kwargs(kwargs: Dict[str, Any], sizes: Sequence[int], batch: int,
                 strict_compat: bool = False) -> List[Dict[str, Any]]:
    n = len(sizes)
    out: List[Dict[str, Any]] = [{} for _ in range(n)]

    def split_any(v: Any) -> Optional[List[Any]]:
        """Return per-worker values, or None to drop the key (compat only)."""
        if _is_batched(v, batch):
            return list(torch.split(v, list(sizes), dim=0))
        if isinstance(v, (list, tuple)) and len(v) > 0 and isinstance(v[0], torch.Tensor):
            if all(_is_batched(t, batch) for t in v):
                cols = [torch.split(t, list(sizes), dim=0) for t in v]
                return [type(v)(c[i] for c in cols) for i in range(n)]
            if strict_compat:
                return None
            parts = [split_any(t) for t in v]
            return [type(v)(p[i] for p in parts) for i in range(n)]
        if isinstance(v, dict) and not strict_compat:
            parts = {k: split_any(x) for k, x in v.items()}
            return [{k: p[i] for k, p in parts.items() if p is not None} for i in range(n)]
        if isinstance(v, (list, tuple)) and not strict_compat and len(v) > 0 and any(
                isinstance(x, (torch.Tensor, dict, list, tuple)) for x in v):
            parts = [split_any(x) for x in v]
            return [type(v)(p[i] for p in parts) for i in range(n)]
        return [v] * n

    for k, v in kwargs.items():
        parts = split_any(v)
        if parts is None:
            continue
        for i in range(n):
            out[i][k] = parts[i]
    return out


def move_to_device(x: Any, device: Any, non_blocking: bool = False, recurse_dicts: bool = True) -> Any:
    """Move tensors (recursively through list/tuple and, unlike ADP:1239-1250, dicts);
    fp8 payloads are widened to fp16 when the target has no fp8 support."""
    device = torch.device(device)
    if isinstance(x, torch.Tensor):
        if x.device != device:
            x = x.to(device, non_blocking=non_blocking)
        if dtypes.is_float8_dtype(x.dtype) and not dtypes.device_supports_float8(device):
            x = x.half()
        return x
    if isinstance(x, (list, tuple)):
        return type(x)(move_to_device(t, device, non_blocking, recurse_dicts) for t in x)
    if isinstance(x, dict) and recurse_dicts:
        return {k: move_to_device(v, device, non_blocking, recurse_dicts) for k, v in x.items()}
    return x


def concatenate_results(results: Sequence[Any], dim: int = 0) -> Any:
    if len(results) == 0:
        return results
    first = results[0]
    if isinstance(first, torch.Tensor):
        return torch.cat(list(results), dim=dim)
    if isinstance(first, (list, tuple)):
        merged = []
        for i, item in enumerate(first):
            if isinstance(item, torch.Tensor):
                merged.append(torch.cat([r[i] for r in results], dim=dim))
            else:
                merged.append(item)
        return type(first)(merged)
    if isinstance(first, dict):
        return {k: concatenate_results([r[k] for r in results], dim) if isinstance(v, (torch.Tensor, list, tuple))
                else v for k, v in first.items()}
    return list(results)


def output_like(first: Any, total: int, device: torch.device) -> Any:
    """Pre-allocate the gathered output given worker-0's output: peers then write
    their rows at final offsets, which removes the ``torch.cat``."""
    if isinstance(first, torch.Tensor):
        return torch.empty((total,) + tuple(first.shape[1:]), dtype=first.dtype, device=device)
    if isinstance(first, (list, tuple)):
        return type(first)(output_like(t, total, device) if isinstance(t, torch.Tensor) else t for t in first)
    return None


def write_rows(dst: Any, src: Any, offset: int) -> None:
    if isinstance(dst, torch.Tensor):
        dst[offset:offset + src.shape[0]].copy_(src, non_blocking=True)
    elif isinstance(dst, (list, tuple)):
        for d, s in zip(dst, src):
            if isinstance(d, torch.Tensor):
                write_rows(d, s, offset)
