"""Config / flag system.

The reference has only ComfyUI widgets (any_device_parallel.py:789-811, 849-865,
886-910) and no env vars or files.  We keep the widgets (see ``nodes.py``) and add
a dataclass + environment overrides used by the bench, tests and debugging
(SURVEY.md §5 "Config / flag system").

Environment variables (all optional):

=====================  ========================================================
``PA_SPLIT_MODE``      ``compat`` (default; reference arithmetic, repaired when
                       the reference would produce an invalid split) | ``exact``
                       (largest-remainder apportionment)
``PA_BACKEND``         ``auto`` | ``fused`` (sm_100a kernels + in-kernel P2P) |
                       ``nccl`` (baseline collectives) | ``torch`` (threads)
``PA_CUDA_GRAPHS``     ``1``/``0`` capture replica forward in CUDA graphs
``PA_FP8``             ``1``: block-scaled fp8 (MXFP8) GEMMs in the native DiT executors
``PA_FAULT``           fault injection, e.g. ``oom:cuda:1@setup``,
                       ``raise:1@step3``, ``oom:1@step2`` (see utils/faults.py)
``PA_FLAG_TIMEOUT_MS`` watchdog for in-kernel flag waits (default 20000)
``PA_LOG_LEVEL``       python logging level
``PA_METRICS_FILE``    JSON-lines metrics sink
``PA_SMALL_BATCH``     ``spread`` (default: 1 < batch < n_devices uses ``batch`` devices) |
                       ``lead`` (reference: lead device only)
``PA_BATCH1``          ``auto`` | ``pipeline`` | ``ulysses`` (batch == 1 mode)
``PA_REPLICATE``       ``auto`` | ``nvls`` | ``p2p`` | ``rebuild``: how further GPUs get the packed weights
``PA_HOST_THREADS``    ``1``/``0`` replay per-GPU CUDA graphs from native host threads
=====================  ========================================================
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field


def _env_bool(name: str, default: bool) -> bool:
    v = os.environ.get(name)
    if v is None:
        return default
    return v.strip().lower() not in ("0", "false", "no", "off", "")


@dataclass
class EngineConfig:
    workload_split: bool = True
    auto_vram_balance: bool = False       # python default of the reference (ADP:917)
    purge_cache: bool = True
    purge_models: bool = False
    split_mode: str = field(default_factory=lambda: os.environ.get("PA_SPLIT_MODE", "compat"))
    backend: str = field(default_factory=lambda: os.environ.get("PA_BACKEND", "auto"))
    cuda_graphs: bool = field(default_factory=lambda: _env_bool("PA_CUDA_GRAPHS", True))
    flag_timeout_ms: int = field(default_factory=lambda: int(os.environ.get("PA_FLAG_TIMEOUT_MS", "20000")))
    fp8: bool = field(default_factory=lambda: _env_bool("PA_FP8", False))   # MXFP8 block GEMMs in native executors
    cache_conditioning: bool = True       # do not re-send constant context every step
    # ComfyUI hands the sampler's cond and uncond halves as one 2B batch; PA_PAIR_CFG=1 keeps sample i and i + B
    # on the same replica (engine._forward_cfg_paired) instead of splitting the 2B rows blindly like the reference
    pair_cfg: bool = field(default_factory=lambda: _env_bool("PA_PAIR_CFG", False))
    # reference semantics for 1 < batch < n_devices is "lead device only" (ADP:1308); by default we instead give one
    # sample to each of the ``batch`` heaviest devices.  ``PA_SMALL_BATCH=lead`` restores the reference behaviour.
    small_batch: str = field(default_factory=lambda: os.environ.get("PA_SMALL_BATCH", "spread"))
    # batch == 1: "pipeline" = the reference's sequential layer split (ADP:1295-1305); "ulysses" = sequence-parallel
    # attention across the chain's GPUs (native FLUX replicas only; falls back to "pipeline" otherwise)
    batch1_mode: str = field(default_factory=lambda: os.environ.get("PA_BATCH1", "auto"))
    # how replicas on further GPUs get their packed weights: "auto" (NVSwitch multicast kernel if the fabric supports
    # it, else peer copies) | "nvls" | "p2p" | "rebuild" (pack again from the torch module on every device)
    replicate: str = field(default_factory=lambda: os.environ.get("PA_REPLICATE", "auto"))
    host_threads: bool = field(default_factory=lambda: _env_bool("PA_HOST_THREADS", True))   # native graph launcher

    def validate(self) -> "EngineConfig":
        if self.split_mode not in ("compat", "exact"):
            raise ValueError(f"split_mode must be compat|exact, got {self.split_mode!r}")
        if self.backend not in ("auto", "fused", "nccl", "torch"):
            raise ValueError(f"backend must be auto|fused|nccl|torch, got {self.backend!r}")
        if self.replicate not in ("auto", "nvls", "p2p", "rebuild"):
            raise ValueError(f"replicate must be auto|nvls|p2p|rebuild, got {self.replicate!r}")
        if self.small_batch not in ("spread", "lead"):
            raise ValueError(f"small_batch must be spread|lead, got {self.small_batch!r}")
        if self.batch1_mode not in ("auto", "pipeline", "ulysses"):
            raise ValueError(f"batch1_mode must be auto|pipeline|ulysses, got {self.batch1_mode!r}")
        return self
