"""Persistent per-device host workers for the in-process (ComfyUI) path.

The reference spins up a fresh ``ThreadPoolExecutor`` on every forward call and
brackets every replica forward with two device-wide synchronisations
(/root/reference/any_device_parallel.py:1385-1391, 1414).  Here each chain slot owns
one long-lived thread (created at setup, joined at cleanup) with its device and
side stream pinned once; ordering between lead and workers is expressed with CUDA
events/stream waits only — the host never blocks on a device.
"""
from __future__ import annotations

import queue
import threading
from concurrent.futures import Future
from typing import Any, Callable, List, Optional

import torch


class device_scope:
    """``with device_scope(dev):`` - make ``dev`` the current device of its backend for the duration of a replica's
    forward and, for backends whose queues we do not order with streams/events (XPU), bracket the work with the
    backend's synchronize - the reference's per-backend branches (/root/reference/any_device_parallel.py:1385-1406:
    cuda -> device + stream, xpu -> ``torch.xpu.device(dev)`` + two syncs, cpu / mps -> plain call)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._ctx = None

    def _backend(self):
        return getattr(torch, self.device.type, None) if self.device.type in ("cuda", "xpu") else None

    def __enter__(self):
        be = self._backend()
        if be is not None and hasattr(be, "device"):
            self._ctx = be.device(self.device)
            self._ctx.__enter__()
            if self.device.type == "xpu" and hasattr(be, "synchronize"):
                be.synchronize(self.device)
        return self

    def __exit__(self, *exc):
        be = self._backend()
        try:
            if self.device.type == "xpu" and be is not None and hasattr(be, "synchronize") and exc[0] is None:
                be.synchronize(self.device)
        finally:
            if self._ctx is not None:
                self._ctx.__exit__(*exc)
        return False


def set_thread_device(device) -> None:
    """Pin a worker thread to its device once (cuda and xpu keep a per-thread current device)."""
    device = torch.device(device)
    be = getattr(torch, device.type, None) if device.type in ("cuda", "xpu") else None
    if be is not None and hasattr(be, "set_device"):
        try:
            be.set_device(device)
        except Exception:
            pass


class DeviceWorker:
    def __init__(self, index: int, device: torch.device, stream: Optional["torch.cuda.Stream"] = None):
        self.index = index
        self.device = torch.device(device)
        self.stream = stream
        self._q: "queue.SimpleQueue" = queue.SimpleQueue()
        self._thread = threading.Thread(target=self._run, name=f"pa-worker-{index}-{device}", daemon=True)
        self._alive = True
        self._thread.start()

    def _run(self) -> None:
        set_thread_device(self.device)
        while True:
            item = self._q.get()
            if item is None:
                return
            fn, fut = item
            if not fut.set_running_or_notify_cancel():
                continue
            try:
                fut.set_result(fn())
            except BaseException as e:  # delivered to the caller, never kills the worker
                fut.set_exception(e)

    def submit(self, fn: Callable[[], Any]) -> Future:
        if not self._alive:
            raise RuntimeError("worker already shut down")
        fut: Future = Future()
        self._q.put((fn, fut))
        return fut

    def shutdown(self, wait: bool = True) -> None:
        if self._alive:
            self._alive = False
            self._q.put(None)
            if wait and threading.current_thread() is not self._thread:
                self._thread.join(timeout=5.0)


class WorkerPool:
    def __init__(self) -> None:
        self.workers: List[DeviceWorker] = []

    def add(self, device: torch.device, stream: Optional["torch.cuda.Stream"] = None) -> DeviceWorker:
        w = DeviceWorker(len(self.workers), device, stream)
        self.workers.append(w)
        return w

    def __len__(self) -> int:
        return len(self.workers)

    def __getitem__(self, i: int) -> DeviceWorker:
        return self.workers[i]

    def shutdown(self) -> None:
        for w in self.workers:
            w.shutdown()
        self.workers = []
