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
        if self.device.type == "cuda":
            try:
                torch.cuda.set_device(self.device)
            except Exception:
                pass
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
