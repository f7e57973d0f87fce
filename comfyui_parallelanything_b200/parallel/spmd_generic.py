"""Multi-process data-parallel forward for an arbitrary ``nn.Module`` replica (any backend).

The family-specific engines (``spmd.py``) move shards inside sm_100a kernels.  This module is the
generic counterpart used (a) for model families without a native executor, (b) as the NCCL baseline
and (c) on CPU with ``gloo`` — which is how the host-side protocol (split, kwargs handling, uneven
shards, empty ranks, error propagation) is tested without GPUs.

Contract: every rank constructs ``SpmdModuleEngine(replica)`` and calls ``forward``; rank 0 passes
the real ``(x, timesteps, context, **kwargs)`` and gets the gathered result, the other ranks call
``forward()`` with no arguments (or run ``serve()``, a loop that exits when rank 0 calls ``stop()``).
Semantics follow the reference's hot path (/root/reference/any_device_parallel.py:1287-1433):
dim-0 split by normalised weights, non-batch kwargs replicated, outputs concatenated on rank 0.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import chain as chain_mod
from . import split as sp

_STOP = "__pa_stop__"


def _flatten(obj: Any, out: List[torch.Tensor]) -> Any:
    """Replace tensors by placeholders, collecting them in order."""
    if isinstance(obj, torch.Tensor):
        out.append(obj)
        return ("__t__", len(out) - 1, tuple(obj.shape), str(obj.dtype).replace("torch.", ""))
    if isinstance(obj, (list, tuple)):
        return type(obj)(_flatten(v, out) for v in obj)
    if isinstance(obj, dict):
        return {k: _flatten(v, out) for k, v in obj.items()}
    return obj


def _is_ph(obj: Any) -> bool:
    return isinstance(obj, tuple) and len(obj) == 4 and obj[0] == "__t__"


def _unflatten(obj: Any, tensors: List[torch.Tensor]) -> Any:
    if _is_ph(obj):
        return tensors[obj[1]]
    if isinstance(obj, (list, tuple)):
        return type(obj)(_unflatten(v, tensors) for v in obj)
    if isinstance(obj, dict):
        return {k: _unflatten(v, tensors) for k, v in obj.items()}
    return obj


def _placeholders(obj: Any, acc: List[tuple]) -> None:
    if _is_ph(obj):
        acc.append(obj)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _placeholders(v, acc)
    elif isinstance(obj, dict):
        for v in obj.values():
            _placeholders(v, acc)


class SpmdModuleEngine:
    def __init__(self, replica: nn.Module, weights: Optional[Sequence[float]] = None, split_mode: str = "compat",
                 group=None, device: Optional[torch.device] = None):
        self.replica = replica
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.weights = chain_mod.normalize_weights(list(weights) if weights is not None else [1.0] * self.world)
        self.split_mode = split_mode
        self.device = device or next((p.device for p in replica.parameters()), torch.device("cpu"))
        self.steps = 0

    # ------------------------------------------------------------------ rank 0
    def forward(self, x=None, timesteps=None, context=None, **kwargs):
        if self.rank != 0:
            return self._worker_once()
        batch = sp.get_batch_size(x)
        if batch < self.world:                       # too small to split: lead only (ADP:1308)
            self._bcast({"mode": "skip"})
            with torch.no_grad():
                return self.replica(x, timesteps, context=context, **kwargs)
        sizes = chain_mod.split_sizes(batch, self.weights, self.split_mode)
        call = {"x": x, "timesteps": timesteps, "context": context, "kwargs": kwargs}
        per_rank = self._split_call(call, sizes, batch)
        metas = []
        payloads: List[List[torch.Tensor]] = []
        for r in range(self.world):
            ts: List[torch.Tensor] = []
            metas.append(_flatten(per_rank[r], ts))
            payloads.append(ts)
        self._bcast({"mode": "run", "sizes": sizes, "metas": metas})
        for r in range(1, self.world):
            if sizes[r] > 0:
                for t in payloads[r]:
                    dist.send(t.contiguous(), dst=r, group=self.group)
        outs: List[Any] = [None] * self.world
        err = None
        if sizes[0] > 0:
            try:
                outs[0] = self._run(per_rank[0])
            except Exception as e:  # noqa: BLE001 - reported after the collective completes
                err = e
        results = self._gather_results(outs[0], sizes, err)
        self.steps += 1
        return results

    def _split_call(self, call: Dict[str, Any], sizes: List[int], batch: int) -> List[Dict[str, Any]]:
        xs = sp.split_value(call["x"], sizes)
        ts = sp.split_value(call["timesteps"], sizes)
        cs = sp.split_value(call["context"], sizes) if call["context"] is not None else [None] * len(sizes)
        ks = sp.split_kwargs(call["kwargs"], sizes, batch)
        return [{"x": xs[i], "timesteps": ts[i], "context": cs[i], "kwargs": ks[i]} for i in range(len(sizes))]

    def _run(self, c: Dict[str, Any]):
        mv = lambda v: sp.move_to_device(v, self.device)  # noqa: E731
        with torch.no_grad():
            return self.replica(mv(c["x"]), mv(c["timesteps"]), context=mv(c["context"]),
                                **{k: mv(v) for k, v in c["kwargs"].items()})

    def _gather_results(self, own: Any, sizes: List[int], err: Optional[BaseException]):
        # every worker first reports (ok, meta-of-output); then sends tensors
        reports: List[Any] = [None] * self.world
        dist.gather_object(("ok", None) if err is None else ("err", repr(err)), reports, dst=0, group=self.group)
        failures = [(r, rep[1]) for r, rep in enumerate(reports) if rep and rep[0] == "err"]
        outs: List[Any] = []
        for r in range(self.world):
            if sizes[r] == 0:
                continue
            if r == 0:
                outs.append(own)
                continue
            if reports[r][0] == "err":
                continue
            meta = reports[r][1]
            phs: List[tuple] = []
            _placeholders(meta, phs)
            ts = []
            for _, _, shape, dt in sorted(phs, key=lambda p: p[1]):
                buf = torch.empty(shape, dtype=getattr(torch, dt), device=self.device)
                dist.recv(buf, src=r, group=self.group)
                ts.append(buf)
            outs.append(_unflatten(meta, ts))
        if failures:
            r, msg = failures[0]
            raise RuntimeError(f"rank {r} failed during the parallel forward: {msg}")
        return sp.concatenate_results(outs, dim=0)

    def stop(self) -> None:
        if self.rank == 0:
            self._bcast({"mode": _STOP})

    # ------------------------------------------------------------------ workers
    def _bcast(self, obj: Any) -> Any:
        box = [obj]
        dist.broadcast_object_list(box, src=0, group=self.group)
        return box[0]

    def _worker_once(self):
        msg = self._bcast(None)
        if msg["mode"] == _STOP:
            return _STOP
        if msg["mode"] == "skip":
            return None
        sizes = msg["sizes"]
        out, err, meta = None, None, None
        if sizes[self.rank] > 0:
            phs: List[tuple] = []
            _placeholders(msg["metas"][self.rank], phs)
            ts = []
            for _, _, shape, dt in sorted(phs, key=lambda p: p[1]):
                buf = torch.empty(shape, dtype=getattr(torch, dt), device=self.device)
                dist.recv(buf, src=0, group=self.group)
                ts.append(buf)
            call = _unflatten(msg["metas"][self.rank], ts)
            try:
                out = self._run(call)
            except Exception as e:  # noqa: BLE001
                err = e
        out_ts: List[torch.Tensor] = []
        if err is None and out is not None:
            meta = _flatten(out, out_ts)
        dist.gather_object(("ok", meta) if err is None else ("err", repr(err)), None, dst=0, group=self.group)
        if err is None:
            for t in out_ts:
                dist.send(t.contiguous(), dst=0, group=self.group)
        self.steps += 1
        return None

    def serve(self) -> int:
        """Worker loop for ranks != 0; returns the number of steps served."""
        n = 0
        while self._worker_once() != _STOP:
            n += 1
        return n
