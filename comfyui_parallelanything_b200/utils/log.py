"""User-visible logging + per-step metrics.

The reference talks to the user through ~49 ``print("[ParallelAnything] ...")``
calls (e.g. /root/reference/any_device_parallel.py:195, 219, 1029, 1467).  We keep
the same prefix so console output stays greppable for people switching over, but
route it through ``logging`` and add a structured metrics sink (JSON lines) that
the reference does not have (SURVEY.md §5 "Metrics / logging").
"""
from __future__ import annotations

import json
import logging
import os
import sys
import threading
import time
from typing import Any, Dict, List, Optional

PREFIX = "[ParallelAnything]"

_logger: Optional[logging.Logger] = None
_lock = threading.Lock()


class _QuietStreamHandler(logging.StreamHandler):
    """stdout may already be closed when weakref finalizers log at interpreter exit."""

    def handleError(self, record):  # noqa: N802
        pass


def get_logger() -> logging.Logger:
    global _logger
    if _logger is None:
        with _lock:
            if _logger is None:
                lg = logging.getLogger("parallel_anything_b200")
                if not lg.handlers:
                    h = _QuietStreamHandler(sys.stdout)
                    h.setFormatter(logging.Formatter(PREFIX + " %(message)s"))
                    lg.addHandler(h)
                    lg.propagate = False
                level = os.environ.get("PA_LOG_LEVEL", "INFO").upper()
                lg.setLevel(getattr(logging, level, logging.INFO))
                _logger = lg
    return _logger


def info(msg: str, *a: Any) -> None:
    get_logger().info(msg, *a)


def warn(msg: str, *a: Any) -> None:
    get_logger().warning("Warning: " + msg, *a)


def error(msg: str, *a: Any) -> None:
    get_logger().error("Error: " + msg, *a)


def debug(msg: str, *a: Any) -> None:
    get_logger().debug(msg, *a)


class Metrics:
    """Tiny counters/timers registry; one per engine.

    ``record(step=..., device_ms=...)`` appends a row; ``dump_jsonl`` writes them.
    If ``PA_METRICS_FILE`` is set every row is also appended there as JSON.
    """

    MAX_ROWS = 4096          # ring: a long-lived ComfyUI session must not grow this without bound

    def __init__(self) -> None:
        self.rows: List[Dict[str, Any]] = []
        self.counters: Dict[str, float] = {}
        self._path = os.environ.get("PA_METRICS_FILE")
        self._lock = threading.Lock()

    def incr(self, key: str, by: float = 1.0) -> None:
        with self._lock:
            self.counters[key] = self.counters.get(key, 0.0) + by

    def record(self, **row: Any) -> None:
        row.setdefault("ts", time.time())
        with self._lock:
            self.rows.append(row)
            if len(self.rows) > self.MAX_ROWS:
                del self.rows[:len(self.rows) - self.MAX_ROWS]
            if self._path:
                try:
                    with open(self._path, "a") as f:
                        f.write(json.dumps(row) + "\n")
                except OSError:
                    pass

    def summary(self) -> Dict[str, Any]:
        with self._lock:
            out: Dict[str, Any] = dict(self.counters)
            ms = [r["device_ms"] for r in self.rows if "device_ms" in r]
            if ms:
                out["steps"] = len(ms)
                out["device_ms_mean"] = sum(ms) / len(ms)
                out["steps_per_sec"] = 1000.0 * len(ms) / sum(ms)
            return out

    def dump_jsonl(self, path: str) -> None:
        with open(path, "w") as f:
            for r in self.rows:
                f.write(json.dumps(r) + "\n")


class nvtx_range:
    """``with nvtx_range("flux.double[3]"):`` — NVTX push/pop when ``PA_NVTX=1`` (for nsys/ncu timelines); a no-op
    otherwise so the hot path pays nothing."""
    enabled = os.environ.get("PA_NVTX", "0") not in ("0", "", "false")

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if self.enabled:
            import torch
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if self.enabled:
            import torch
            torch.cuda.nvtx.range_pop()
        return False
