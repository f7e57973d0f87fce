"""ComfyUI node API — schema-identical to the reference so saved graphs keep working.

Parity targets in /root/reference/any_device_parallel.py:
  * ``ParallelDevice``      ADP:768-832   ("Parallel Device Config")
  * ``ParallelDeviceList``  ADP:834-882   ("Parallel Device List (1-4x)")
  * ``ParallelAnything``    ADP:884-1471  ("Parallel Anything (True Multi-GPU)")
  * registration            ADP:1473-1483

Field names, defaults, min/max/step, RETURN_TYPES/NAMES, FUNCTION, CATEGORY and the
mapping keys are kept; the implementation behind ``setup_parallel`` is the new
engine (``engine.py``), which on B200s swaps torch replicas for hand-written
sm_100a executors when the wrapped model family is known.
"""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import torch
import torch.nn as nn

from . import chain as chain_mod
from . import engine as engine_mod
from .utils import log, memory
from .utils.config import EngineConfig


def _discover(include_directml: bool) -> List[str]:
    devices = ["cpu"]
    if torch.cuda.is_available():
        devices += [f"cuda:{i}" for i in range(torch.cuda.device_count())]
    mps = getattr(torch.backends, "mps", None)
    if mps is not None and mps.is_available():
        devices.append("mps")
    xpu = getattr(torch, "xpu", None)
    if xpu is not None and xpu.is_available():
        devices += [f"xpu:{i}" for i in range(xpu.device_count())]
    if include_directml:
        try:
            import torch_directml  # type: ignore
            devices += [f"privateuseone:{i}" for i in range(torch_directml.device_count())]
        except ImportError:
            pass
    return devices


class ParallelDevice:
    @classmethod
    def get_available_devices(cls) -> List[str]:
        return _discover(include_directml=True)

    @classmethod
    def INPUT_TYPES(cls):
        available = cls.get_available_devices()
        default = "cuda:0" if any(d.startswith("cuda:0") for d in available) else available[0]
        return {
            "required": {
                "device_id": (available, {
                    "default": default,
                    "tooltip": "Select available compute device (CPU/CUDA/MPS/XPU)",
                }),
                "percentage": ("FLOAT", {
                    "default": 50.0, "min": 1.0, "max": 100.0, "step": 1.0,
                    "tooltip": "Percentage of batch (or layers for batch=1) to process on this device",
                }),
            },
            "optional": {
                "previous_devices": ("DEVICE_CHAIN", {
                    "tooltip": "Connect from another ParallelDevice node to chain multiple GPUs",
                }),
            },
        }

    RETURN_TYPES = ("DEVICE_CHAIN",)
    RETURN_NAMES = ("device_chain",)
    FUNCTION = "add_device"
    CATEGORY = "utils/hardware"
    DESCRIPTION = "Add a GPU/CPU/MPS/XPU device to the parallel processing chain"

    def add_device(self, device_id, percentage, previous_devices=None):
        chain = list(previous_devices) if previous_devices else []
        chain.append(chain_mod.make_entry(device_id, percentage))
        return (chain,)


class ParallelDeviceList:
    @classmethod
    def get_available_devices(cls) -> List[str]:
        return _discover(include_directml=False)

    @classmethod
    def INPUT_TYPES(cls):
        devices = cls.get_available_devices()
        first = "cuda:0" if "cuda:0" in devices else devices[0]

        def nth(i: int, fallback: str) -> str:
            return devices[i] if len(devices) > i else fallback

        pct = lambda default, lo: ("FLOAT", {"default": default, "min": lo, "max": 100.0, "step": 1.0})
        return {
            "required": {
                "device_1": (devices, {"default": first}),
                "pct_1": pct(50.0, 1.0),
                "device_2": (devices, {"default": nth(1, first)}),
                "pct_2": pct(50.0, 0.0),
            },
            "optional": {
                "device_3": (devices, {"default": nth(2, "cpu")}),
                "pct_3": pct(0.0, 0.0),
                "device_4": (devices, {"default": nth(3, "cpu")}),
                "pct_4": pct(0.0, 0.0),
            },
        }

    RETURN_TYPES = ("DEVICE_CHAIN",)
    RETURN_NAMES = ("device_chain",)
    FUNCTION = "create_list"
    CATEGORY = "utils/hardware"

    def create_list(self, device_1, pct_1, device_2, pct_2, device_3="cpu", pct_3=0, device_4="cpu", pct_4=0):
        pairs = ((device_1, pct_1), (device_2, pct_2), (device_3, pct_3), (device_4, pct_4))
        return ([chain_mod.make_entry(d, p) for d, p in pairs if p > 0],)


# ------------------------------------------------------------------- MODEL plumbing

def unwrap_model(model: Any) -> nn.Module:
    """ModelPatcher -> BaseModel -> diffusion_model; BaseModel -> diffusion_model; or a
    bare nn.Module (ADP:922-930)."""
    inner = getattr(model, "model", None)
    if inner is not None and hasattr(inner, "diffusion_model"):
        return inner.diffusion_model
    if hasattr(model, "diffusion_model"):
        return model.diffusion_model
    return model


def _find_patches(model: Any) -> Tuple[bool, Any]:
    """Non-empty LoRA ``patches`` on the patcher (three places, ADP:977-981)."""
    for holder in (model, getattr(model, "model", None), getattr(model, "patcher", None)):
        patches = getattr(holder, "patches", None) if holder is not None else None
        if patches:
            try:
                if len(patches) > 0:
                    return True, holder
            except TypeError:
                return True, holder
    return False, None


def _repair_stranded(model: Any, target: nn.Module) -> None:
    """A previous (reference-style) run may have left the weights on CPU while ComfyUI
    still believes they are on the GPU (ADP:932-961).  We never strand the model
    ourselves, but repair it when we find it that way."""
    cur = memory.module_device(target)
    if cur is None or cur.type != "cpu":
        return
    want = getattr(model, "load_device", None)
    if want is None and getattr(model, "model", None) is not None:
        want = getattr(model.model, "load_device", None)
    if want is None:
        mm = memory.comfy_mm()
        if mm is not None:
            try:
                want = mm.get_torch_device()
            except Exception:
                want = None
    if want is None:
        return
    want = torch.device(want)
    if want.type == "cuda" and torch.cuda.is_available():
        log.info("Model was stranded on CPU; moving it back to %s", want)
        try:
            target.to(want)
        except Exception as e:
            log.warn("could not move the model back to %s: %s", want, e)


class ParallelAnything:
    @classmethod
    def INPUT_TYPES(cls):
        return {
            "required": {
                "model": ("MODEL",),
                "device_chain": ("DEVICE_CHAIN", {"tooltip": "Connect from ParallelDevice nodes"}),
            },
            "optional": {
                "workload_split": ("BOOLEAN", {"default": True, "tooltip": "Enable multi-device processing"}),
                "auto_vram_balance": ("BOOLEAN", {
                    "default": True,
                    "tooltip": "Automatically adjust batch split based on available VRAM"}),
                "purge_cache": ("BOOLEAN", {
                    "default": True,
                    "tooltip": "Purge CUDA cache when cleaning up parallel resources"}),
                "purge_models": ("BOOLEAN", {
                    "default": False,
                    "tooltip": "Unload all models from VRAM when cleaning up (aggressive memory clearing)"}),
            },
        }

    RETURN_TYPES = ("MODEL",)
    RETURN_NAMES = ("model",)
    FUNCTION = "setup_parallel"
    CATEGORY = "utils/hardware"

    def setup_parallel(self, model, device_chain, workload_split=True, auto_vram_balance=False,
                       purge_cache=True, purge_models=False, config: Optional[EngineConfig] = None):
        """Returns ``(model,)`` — the *same* object, mutated in place (ADP:1471)."""
        if model is None or not device_chain:
            return (model,)
        target = unwrap_model(model)
        if not isinstance(target, nn.Module):
            log.error("MODEL does not contain an nn.Module diffusion model; leaving it untouched")
            return (model,)

        _repair_stranded(model, target)
        original_device = memory.module_device(target) or torch.device("cpu")

        has_lora, holder = _find_patches(model)
        if has_lora and hasattr(holder, "patch_model"):
            mm = memory.comfy_mm()
            try:
                dev_to = mm.get_torch_device() if mm is not None else original_device
                log.info("LoRA patches detected: baking them into the weights before cloning")
                holder.patch_model(device_to=dev_to)
                original_device = memory.module_device(target) or original_device
            except Exception as e:
                log.warn("could not apply LoRA patches before cloning: %s", e)

        if getattr(target, "_true_parallel_active", False):
            log.info("Model already parallelised; tearing down the previous setup")
            engine_mod.cleanup_parallel_model(target)
        mm = memory.comfy_mm()
        if mm is not None:
            try:
                mm.unload_all_models()
            except Exception:
                pass
            memory.aggressive_cleanup()

        cfg = config or EngineConfig()
        cfg.workload_split = bool(workload_split)
        cfg.auto_vram_balance = bool(auto_vram_balance)
        cfg.purge_cache = bool(purge_cache)
        cfg.purge_models = bool(purge_models)

        eng = engine_mod.ParallelEngine(target, device_chain, cfg)
        try:
            ok = eng.setup(has_lora=has_lora, original_device=original_device)
        except Exception as e:  # never break the graph: hand back the untouched model
            log.error("Parallel setup raised %s: %s", type(e).__name__, e)
            eng.cleanup()
            ok = False
        if not ok:
            return (model,)

        engine_mod.install(eng, owner=model)

        lead = eng.lead_device
        if hasattr(model, "load_device"):
            try:
                model.load_device = lead
            except Exception:
                pass
        elif getattr(model, "model", None) is not None and hasattr(model.model, "load_device"):
            model.model.load_device = lead

        log.info("Parallel setup complete. Devices: %s", eng.device_names)
        if has_lora:
            log.info("LoRA weights synchronized across all devices")
        return (model,)


NODE_CLASS_MAPPINGS = {
    "ParallelAnything": ParallelAnything,
    "ParallelDevice": ParallelDevice,
    "ParallelDeviceList": ParallelDeviceList,
}

NODE_DISPLAY_NAME_MAPPINGS = {
    "ParallelAnything": "Parallel Anything (True Multi-GPU)",
    "ParallelDevice": "Parallel Device Config",
    "ParallelDeviceList": "Parallel Device List (1-4x)",
}
