"""Batch==1 pipeline ("layer-split") mode.

Reference: ``ParallelBlock`` /root/reference/any_device_parallel.py:24-87, plan
ADP:1152-1198, thread-local switch ADP:16-22.

Differences by design:
  * the stage wrapper does **not** register the other replicas' blocks as
    submodules (the reference's attribute assignment makes the lead replica's
    ``.parameters()/.to()`` reach into foreign replicas) — peers are held
    in a plain tuple behind ``object.__setattr__``;
  * hand-offs only happen at ownership *boundaries*: a block whose inputs already
    live on the owner device costs no copies (the reference re-walks every arg of
    every block);
  * wrappers are removable (``unwrap_blocks``) so a second setup never nests
    wrapper-in-wrapper (Appendix A16 [PROBE]);
  * WAN's ``blocks`` list is included (Appendix A18).
"""
from __future__ import annotations

import copy
import dataclasses
import threading
from typing import Any, Dict, List, Sequence

import torch
import torch.nn as nn

from .. import chain as chain_mod
from ..utils import log

BLOCK_LIST_NAMES = ("double_blocks", "single_blocks", "transformer_blocks", "layers", "blocks")

_state = threading.local()


def get_pipeline_mode() -> bool:
    return getattr(_state, "active", False)


def set_pipeline_mode(active: bool) -> None:
    _state.active = bool(active)


class pipeline_mode:
    """Context manager; thread-local so DP worker threads never take the PP path."""

    def __init__(self, active: bool = True):
        self.active = active

    def __enter__(self):
        self.prev = get_pipeline_mode()
        set_pipeline_mode(self.active)
        return self

    def __exit__(self, *exc):
        set_pipeline_mode(self.prev)
        return False


def move_nested(x: Any, device: torch.device) -> Any:
    if isinstance(x, torch.Tensor):
        return x if x.device == device else x.to(device, non_blocking=True)
    if isinstance(x, (list, tuple)):
        return type(x)(move_nested(v, device) for v in x)
    if isinstance(x, dict):
        return {k: move_nested(v, device) for k, v in x.items()}
    if dataclasses.is_dataclass(x) and not isinstance(x, type):
        new = copy.copy(x)
        for f in dataclasses.fields(x):
            object.__setattr__(new, f.name, move_nested(getattr(x, f.name), device))
        return new
    return x


class PipelineStage(nn.Module):
    """Drop-in replacement for one transformer block of the lead replica.

    DP mode (default): run ``local_block``.  Pipeline mode: run the same-index block
    of the replica living on ``owner_device``; move inputs there first and, for the
    last block of the list, move the result back to ``lead_device``."""

    def __init__(self, local_block: nn.Module, block_idx: int, owner_device: torch.device,
                 owner_block: nn.Module, is_last_block: bool, lead_device: torch.device):
        super().__init__()
        self.local_block = local_block                       # registered: it *is* ours
        self.block_idx = block_idx
        self.owner_device = torch.device(owner_device)
        self.lead_device = torch.device(lead_device)
        self.is_last_block = bool(is_last_block)
        object.__setattr__(self, "_owner", (owner_block,))   # NOT registered

    @property
    def owner_block(self) -> nn.Module:
        return self._owner[0]

    def forward(self, *args, **kwargs):
        if not get_pipeline_mode():
            return self.local_block(*args, **kwargs)
        dev = self.owner_device
        args = tuple(move_nested(a, dev) for a in args)
        kwargs = {k: move_nested(v, dev) for k, v in kwargs.items()}
        if dev.type == "cuda":
            with torch.cuda.device(dev):
                out = self.owner_block(*args, **kwargs)
        else:
            out = self.owner_block(*args, **kwargs)
        if self.is_last_block:
            out = move_nested(out, self.lead_device)
        return out


def unwrap_blocks(model: nn.Module) -> int:
    """Replace every PipelineStage in the known block lists by its local block."""
    n = 0
    for name in BLOCK_LIST_NAMES:
        blocks = getattr(model, name, None)
        if not isinstance(blocks, nn.ModuleList):
            continue
        for i, b in enumerate(blocks):
            while isinstance(b, PipelineStage):
                b = b.local_block
                n += 1
            blocks[i] = b
    return n


def plan(model: nn.Module, weights: Sequence[float]) -> Dict[str, List[int]]:
    """owner index per block for every eligible block list of ``model``."""
    out: Dict[str, List[int]] = {}
    for name in BLOCK_LIST_NAMES:
        blocks = getattr(model, name, None)
        if isinstance(blocks, nn.ModuleList) and len(blocks) > 0:
            out[name] = chain_mod.assign_blocks(len(blocks), weights)
    return out


def wrap_blocks(lead_replica: nn.Module, replicas: Dict[str, nn.Module], device_names: Sequence[str],
                weights: Sequence[float]) -> Dict[str, List[int]]:
    """Install PipelineStage wrappers in ``lead_replica`` according to ``plan``."""
    unwrap_blocks(lead_replica)
    assignment = plan(lead_replica, weights)
    lead_dev = torch.device(device_names[0])
    for name, owners in assignment.items():
        local_blocks = getattr(lead_replica, name)
        log.info("Configuring %s (%d blocks) for pipeline execution: %s", name, len(owners),
                 _summarise(owners, device_names))
        for idx, owner_idx in enumerate(owners):
            dev_name = device_names[owner_idx]
            owner_replica = replicas[dev_name]
            peer_list = getattr(owner_replica, name, None)
            owner_block = local_blocks[idx]
            owner_dev = lead_dev
            if owner_replica is lead_replica:
                owner_dev = torch.device(dev_name)
            elif isinstance(peer_list, nn.ModuleList) and idx < len(peer_list):
                cand = peer_list[idx]
                while isinstance(cand, PipelineStage):
                    cand = cand.local_block
                owner_block, owner_dev = cand, torch.device(dev_name)
            # else: the owner replica exposes no block list (a native executor holds packed weights, not modules):
            # the stage stays on the lead device with the lead's own block - never a lead block fed foreign tensors.
            local_blocks[idx] = PipelineStage(local_blocks[idx], idx, owner_dev, owner_block,
                                              idx == len(owners) - 1, lead_dev)
    return assignment


def _summarise(owners: Sequence[int], device_names: Sequence[str]) -> str:
    parts, start = [], 0
    for i in range(1, len(owners) + 1):
        if i == len(owners) or owners[i] != owners[start]:
            parts.append(f"{device_names[owners[start]]}[{start}:{i}]")
            start = i
    return " ".join(parts)
