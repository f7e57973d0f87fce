"""dtype / device-capability predicates.

Parity targets (behaviour, not code): ``is_float8_dtype`` ADP:93-98,
``check_sm80_support`` ADP:100-110, ``device_supports_float8`` ADP:112-124 in
/root/reference/any_device_parallel.py.  On B200 (sm_100) fp8 is native, so the
"downcast fp8 -> fp16" rule (ADP:403-404, 654-655, 1243-1244) only ever triggers
for non-CUDA / pre-Hopper devices in the torch fallback path.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch

_FP8_DTYPES = tuple(
    getattr(torch, n)
    for n in ("float8_e4m3fn", "float8_e5m2", "float8_e4m3fnuz", "float8_e5m2fnuz", "float8_e8m0fnu")
    if hasattr(torch, n)
)


def is_float8_dtype(dtype: Optional[torch.dtype]) -> bool:
    if dtype is None:
        return False
    if dtype in _FP8_DTYPES:
        return True
    return "float8" in str(dtype)  # future fp8 flavours


def _cuda_index(device: Union[str, torch.device]) -> Optional[int]:
    d = torch.device(device) if not isinstance(device, torch.device) else device
    if d.type != "cuda":
        return None
    return d.index if d.index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)


def cuda_capability(device: Union[str, torch.device]) -> Optional[Tuple[int, int]]:
    idx = _cuda_index(device)
    if idx is None:
        return None
    try:
        return tuple(torch.cuda.get_device_capability(idx))  # type: ignore[return-value]
    except Exception:
        return None


def check_sm80_support(device: Union[str, torch.device]) -> bool:
    """Non-CUDA devices pass; CUDA devices need major >= 8.  Unknown -> False."""
    try:
        d = torch.device(device)
    except Exception:
        return False
    if d.type != "cuda":
        return True
    cap = cuda_capability(d)
    return cap is not None and cap[0] >= 8


def device_supports_float8(device: Union[str, torch.device]) -> bool:
    """CUDA and capability >= (9, 0) — Ada (8.9) is deliberately excluded, matching
    the reference's rule; B200 is (10, 0)."""
    try:
        d = torch.device(device)
    except Exception:
        return False
    cap = cuda_capability(d)
    return cap is not None and cap >= (9, 0)


def is_blackwell(device: Union[str, torch.device]) -> bool:
    cap = cuda_capability(device)
    return cap is not None and cap[0] == 10


def storage_dtype_for(dtype: torch.dtype, device: Union[str, torch.device]) -> torch.dtype:
    """dtype a tensor of ``dtype`` should have once it lives on ``device``."""
    if is_float8_dtype(dtype) and not device_supports_float8(device):
        return torch.float16
    return dtype
