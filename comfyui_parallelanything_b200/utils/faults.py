"""Env-driven fault injection (SURVEY.md §5 "Failure detection ... fault injection").

The reference has OOM-centric recovery only (skip device at clone time
ADP:1114-1128, lead-only rerun at forward time ADP:1435-1446) and no way to test
it.  ``PA_FAULT`` lets tests trigger those paths deterministically:

    PA_FAULT="oom:cuda:1@setup"       raise a CUDA-OOM-looking error while cloning onto cuda:1
    PA_FAULT="oom:1@step2"            worker index 1 raises OOM on forward call #2
    PA_FAULT="raise:0@step0"          worker index 0 raises a generic RuntimeError
    PA_FAULT="hang:1@step1"           worker 1 sleeps past the watchdog

Several faults may be joined with ``;``.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass
from typing import List, Optional


class InjectedOOM(RuntimeError):
    def __init__(self, where: str):
        super().__init__(f"CUDA out of memory (injected fault at {where})")


@dataclass
class Fault:
    kind: str      # oom | raise | hang
    target: str    # device string or worker index
    when: str      # "setup" or "step<N>"


def parse(spec: Optional[str] = None) -> List[Fault]:
    spec = os.environ.get("PA_FAULT", "") if spec is None else spec
    out: List[Fault] = []
    for part in filter(None, (s.strip() for s in spec.split(";"))):
        try:
            left, when = part.rsplit("@", 1)
            kind, target = left.split(":", 1)
            out.append(Fault(kind.strip(), target.strip(), when.strip()))
        except ValueError:
            raise ValueError(f"bad PA_FAULT entry {part!r}")
    return out


def _matches(f: Fault, device: str, index: Optional[int]) -> bool:
    return f.target == device or (index is not None and f.target == str(index))


def check_setup(device: str, index: Optional[int] = None) -> None:
    for f in parse():
        if f.when == "setup" and _matches(f, device, index):
            if f.kind == "oom":
                raise InjectedOOM(f"setup:{device}")
            raise RuntimeError(f"injected fault at setup:{device}")


def check_step(step: int, device: str, index: Optional[int] = None) -> None:
    for f in parse():
        if f.when == f"step{step}" and _matches(f, device, index):
            if f.kind == "oom":
                raise InjectedOOM(f"step{step}:{device}")
            if f.kind == "hang":
                time.sleep(float(os.environ.get("PA_FAULT_HANG_S", "5")))
                return
            raise RuntimeError(f"injected fault at step{step}:{device}")
