"""DEVICE_CHAIN data model, weight normalisation and batch apportionment.

``DEVICE_CHAIN`` is the custom ComfyUI socket type of the reference: a plain
``list[dict(device:str, percentage:float, weight:float)]``
(/root/reference/any_device_parallel.py:823-832, 872-882).  We keep that wire
format (so graphs saved for the reference load unchanged) and add typed helpers.

Apportionment:
  * ``compat``  — the reference arithmetic ``max(1, int(B*w))`` with the last entry
    taking the remainder (ADP:1321-1322).  The reference can produce a *negative*
    last entry (B=4, weights .9/.03/.03/.04 -> [3,1,1,-1]) and then crashes in
    ``torch.split``; in that case we fall through to ``exact``.
  * ``exact``   — largest-remainder (Hamilton) apportionment, every entry >= 0,
    always sums to B.
  * ``vram``    — the reference's 70/30 blend of weight and free-VRAM share
    (ADP:737-766), fixed so that it always conserves B.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch

from .utils import memory


@dataclass(frozen=True)
class DeviceEntry:
    device: str
    percentage: float

    @property
    def weight(self) -> float:
        return self.percentage / 100.0

    def as_dict(self) -> Dict[str, float]:
        return {"device": self.device, "percentage": float(self.percentage), "weight": self.weight}


def make_entry(device: str, percentage: float) -> Dict[str, float]:
    return DeviceEntry(str(device), float(percentage)).as_dict()


def parse_chain(chain: Iterable) -> List[DeviceEntry]:
    """Accept the reference wire format, DeviceEntry objects or (device, pct) pairs."""
    out: List[DeviceEntry] = []
    for item in chain or []:
        if isinstance(item, DeviceEntry):
            out.append(item)
        elif isinstance(item, dict):
            pct = item.get("percentage")
            if pct is None:
                pct = float(item.get("weight", 0.0)) * 100.0
            out.append(DeviceEntry(str(item["device"]), float(pct)))
        else:
            dev, pct = item
            out.append(DeviceEntry(str(dev), float(pct)))
    return out


def normalize_weights(percentages: Sequence[float]) -> List[float]:
    """``w_i = pct_i / sum(pct)`` — percentages need not total 100 (ADP:1019-1027).
    A non-positive total degrades to an even split."""
    total = float(sum(percentages))
    n = len(percentages)
    if n == 0:
        return []
    if total <= 0:
        return [1.0 / n] * n
    return [float(p) / total for p in percentages]


def validate_devices(names: Sequence[str]) -> Optional[str]:
    """Return the first invalid device string or None (ADP:1037-1042)."""
    for n in names:
        try:
            d = torch.device(n)
        except Exception:
            return n
        if d.type == "cuda":
            if not torch.cuda.is_available():
                return n
            if d.index is not None and d.index >= torch.cuda.device_count():
                return n
    return None


# --------------------------------------------------------------------------- splits

def split_compat(batch: int, weights: Sequence[float]) -> List[int]:
    sizes = [max(1, int(batch * w)) for w in weights]
    sizes[-1] = batch - sum(sizes[:-1])
    return sizes


def split_exact(batch: int, weights: Sequence[float]) -> List[int]:
    w = normalize_weights(list(weights))
    quotas = [batch * x for x in w]
    sizes = [int(q) for q in quotas]
    rest = batch - sum(sizes)
    order = sorted(range(len(w)), key=lambda i: (-(quotas[i] - sizes[i]), i))
    for i in order[:rest]:
        sizes[i] += 1
    return sizes


def _valid(sizes: Sequence[int], batch: int) -> bool:
    return all(s >= 0 for s in sizes) and sum(sizes) == batch


def split_sizes(batch: int, weights: Sequence[float], mode: str = "compat") -> List[int]:
    if not weights:
        return []
    if mode == "compat":
        s = split_compat(batch, weights)
        if _valid(s, batch):
            return s
    return split_exact(batch, weights)


def vram_adjusted_weights(devices: Sequence[str], weights: Sequence[float],
                          free_vram: Optional[Callable[[str], float]] = None) -> List[float]:
    """0.7*w + 0.3*(free_i / sum free) for CUDA devices with non-zero free memory."""
    free_vram = free_vram or memory.get_free_vram
    avail = [free_vram(d) if str(d).startswith("cuda") else 0.0 for d in devices]
    total = sum(avail)
    if total <= 0:
        return list(weights)
    adj = [0.7 * w + 0.3 * (v / total) if v > 0 else w for w, v in zip(weights, avail)]
    return normalize_weights(adj)


def split_sizes_vram(batch: int, devices: Sequence[str], weights: Sequence[float],
                     mode: str = "compat",
                     free_vram: Optional[Callable[[str], float]] = None) -> List[int]:
    return split_sizes(batch, vram_adjusted_weights(devices, weights, free_vram), mode)


def assign_blocks(num_blocks: int, weights: Sequence[float]) -> List[int]:
    """Pipeline plan: contiguous block ranges, ``round(w*num_blocks)`` each, the last
    device takes the remainder (ADP:1168-1178).  Returns owner index per block."""
    owners: List[int] = []
    cur = 0
    n = len(weights)
    for i, w in enumerate(weights):
        count = int(round(w * num_blocks))
        if i == n - 1:
            count = num_blocks - cur
        count = max(0, min(count, num_blocks - cur))
        owners.extend([i] * count)
        cur += count
    if len(owners) < num_blocks:  # rounding starved the tail: give the rest to the last device
        owners.extend([n - 1] * (num_blocks - len(owners)))
    return owners


def offsets(sizes: Sequence[int]) -> List[int]:
    out, acc = [], 0
    for s in sizes:
        out.append(acc)
        acc += s
    return out
