"""Model replication onto another device.

What the reference does (/root/reference/any_device_parallel.py):
  * ``extract_model_config``      ADP:284-350  harvest ctor kwargs
  * ``clone_dataclass_or_object`` ADP:352-385
  * ``clone_module_simple``       ADP:390-584  structural fallback clone
  * ``safe_model_clone``          ADP:586-722  source.cpu() *in place*, rebuild from
    config, per-key H2D state_dict load, fp8->fp16 on pre-sm_90 targets, freeze.

B200-first redesign: replicas are produced **device-to-device** — every parameter
and buffer is copied with one ``Tensor.to(target)`` (NVLink P2P when both ends are
CUDA) while the module *structure* is duplicated by ``copy.deepcopy`` with a
pre-seeded memo, so no weight ever bounces through host memory, the source is
never moved (fixes Appendix A11 "stranded on CPU"), and peak extra memory on the
target is exactly one replica.  The reference's two strategies are kept as
fallbacks (rebuild-from-config, then a structural clone) for exotic modules whose
``__deepcopy__`` fails.
"""
from __future__ import annotations

import copy
import dataclasses
from types import SimpleNamespace
from typing import Any, Dict, Optional, Union

import torch
import torch.nn as nn

from . import dtypes, faults, log, memory

DeviceLike = Union[str, torch.device]

# attribute names that commonly carry constructor arguments in ComfyUI model classes
_CONFIG_ATTRS = (
    # flux-like DiT
    "in_channels", "out_channels", "vec_in_dim", "context_in_dim", "hidden_size", "mlp_ratio",
    "num_heads", "depth", "depth_single_blocks", "depth_single", "axes_dim", "theta", "patch_size",
    "qkv_bias", "guidance_embed", "txt_ids_dim", "img_ids_dim",
    # ldm UNet
    "num_res_blocks", "attention_resolutions", "dropout", "channel_mult", "num_classes",
    "use_checkpoint", "num_heads_upsample", "use_scale_shift_norm", "resblock_updown",
    "use_new_attention_order", "adm_in_channels", "num_noises", "context_dim", "n_heads", "d_head",
    "transformer_depth", "model_channels", "max_depth",
    # video
    "num_frames", "temporal_compression", "temporal_dim", "video_length",
)

_SKIP_TYPES = (torch.Tensor, nn.Module)


def _plain(v: Any) -> bool:
    return v is not None and not callable(v) and not isinstance(v, _SKIP_TYPES)


def extract_model_config(model: nn.Module) -> Dict[str, Any]:
    """Best-effort constructor kwargs of ``model``: well-known attributes, then
    ``model.params`` (dict / dataclass / object), ``model.config`` and
    ``model.unet_config``; only deep-copyable, tensor-free values survive."""
    cfg: Dict[str, Any] = {}
    for name in _CONFIG_ATTRS:
        try:
            v = getattr(model, name)
        except Exception:
            continue
        if _plain(v):
            cfg[name] = v

    def merge(obj: Any, skip_private: bool = False) -> None:
        if obj is None:
            return
        if isinstance(obj, dict):
            items = obj.items()
        elif dataclasses.is_dataclass(obj) and not isinstance(obj, type):
            items = ((f.name, getattr(obj, f.name, None)) for f in dataclasses.fields(obj))
        else:
            try:
                items = vars(obj).items()
            except TypeError:
                return
        for k, v in items:
            if skip_private and str(k).startswith("_"):
                continue
            if not isinstance(v, _SKIP_TYPES) and not (skip_private and callable(v)):
                cfg[k] = v

    for attr, private in (("params", False), ("config", True)):
        try:
            merge(getattr(model, attr, None), skip_private=private)
        except Exception:
            pass
    uc = getattr(model, "unet_config", None)
    if isinstance(uc, dict):
        merge(uc)

    clean: Dict[str, Any] = {}
    for k, v in cfg.items():
        if v is None or isinstance(v, _SKIP_TYPES):
            continue
        try:
            copy.deepcopy(v)
        except Exception:
            continue
        clean[k] = v
    return clean


def clone_dataclass_or_object(obj: Any, target_device: Optional[DeviceLike] = None) -> Any:
    """Field-wise clone of dataclasses (tensors detached+cloned, optionally moved),
    deep copy for everything else, the object itself as a last resort."""

    def one(v: Any) -> Any:
        if isinstance(v, torch.Tensor):
            t = v.detach().clone()
            return t.to(target_device) if target_device is not None else t
        if dataclasses.is_dataclass(v) and not isinstance(v, type):
            return clone_dataclass_or_object(v, target_device)
        if isinstance(v, (list, tuple)):
            return type(v)(one(x) for x in v)
        return copy.deepcopy(v)

    if dataclasses.is_dataclass(obj) and not isinstance(obj, type):
        try:
            vals = {}
            for f in dataclasses.fields(obj):
                cur = getattr(obj, f.name)
                try:
                    vals[f.name] = one(cur)
                except Exception:
                    vals[f.name] = cur
            return dataclasses.replace(obj, **{k: v for k, v in vals.items() if _is_init_field(obj, k)})
        except Exception:
            pass
    try:
        return copy.deepcopy(obj)
    except Exception:
        return obj


def _is_init_field(obj: Any, name: str) -> bool:
    for f in dataclasses.fields(obj):
        if f.name == name:
            return f.init
    return False


# --------------------------------------------------------------------------- tensor moves

def _move_tensor(t: torch.Tensor, device: torch.device, keep_grad: bool = False) -> torch.Tensor:
    dt = dtypes.storage_dtype_for(t.dtype, device)
    out = t.detach().to(device=device, dtype=dt, copy=True)
    if keep_grad and t.requires_grad:
        out.requires_grad_(True)
    return out


def clone_module_d2d(module: nn.Module, device: DeviceLike) -> nn.Module:
    """deepcopy the module graph while every Parameter/buffer is materialised
    directly on ``device`` (one D2D copy each, shared tensors stay shared)."""
    device = torch.device(device)
    memo: Dict[int, Any] = {}
    for p in module.parameters():
        if id(p) not in memo:
            memo[id(p)] = nn.Parameter(_move_tensor(p, device), requires_grad=False)
    for b in module.buffers():
        if id(b) not in memo:
            memo[id(b)] = _move_tensor(b, device)
    clone = copy.deepcopy(module, memo)
    return clone


def clone_module_structural(module: Optional[nn.Module], device: DeviceLike) -> Optional[nn.Module]:
    """Fallback clone that never calls ``__init__``/``__deepcopy__`` of custom classes:
    allocate with ``__new__``, re-register parameters/buffers on ``device``, recurse
    into children, then carry over the remaining instance attributes (tensors moved,
    dataclasses cloned, cache-like attributes dropped).  Raises RuntimeError on failure."""
    if module is None:
        return None
    device = torch.device(device)
    try:
        cls = module.__class__
        new = cls.__new__(cls)
        nn.Module.__init__(new)
        for name, p in module._parameters.items():
            new._parameters[name] = None if p is None else nn.Parameter(_move_tensor(p, device), requires_grad=False)
        for name, b in module._buffers.items():
            new._buffers[name] = None if b is None else _move_tensor(b, device)
        new._non_persistent_buffers_set = set(getattr(module, "_non_persistent_buffers_set", ()))
        for name, child in module._modules.items():
            new._modules[name] = clone_module_structural(child, device)
        reserved = set(nn.Module().__dict__.keys())
        for k, v in module.__dict__.items():
            if k in reserved:
                continue
            if k in memory.CACHE_ATTRS:
                new.__dict__[k] = None
            elif isinstance(v, torch.Tensor):
                new.__dict__[k] = _move_tensor(v, device)
            elif isinstance(v, nn.Module):
                new.__dict__[k] = clone_module_structural(v, device)
            elif dataclasses.is_dataclass(v) and not isinstance(v, type):
                new.__dict__[k] = clone_dataclass_or_object(v, device)
            else:
                try:
                    new.__dict__[k] = copy.deepcopy(v)
                except Exception:
                    new.__dict__[k] = v
        new.training = module.training
        return new
    except Exception as e:  # pragma: no cover - defensive
        raise RuntimeError(f"structural clone of {type(module).__name__} failed: {e}") from e


def clone_module_from_config(module: nn.Module, device: DeviceLike) -> nn.Module:
    """Reference-style rebuild: ``cls(**config)`` / ``cls(config)`` /
    ``cls(SimpleNamespace(**config))`` then a key-by-key state load on the target."""
    device = torch.device(device)
    cfg = extract_model_config(module)
    cls = module.__class__
    new = None
    errors = []
    for build in (lambda: cls(**cfg), lambda: cls(cfg), lambda: cls(SimpleNamespace(**cfg))):
        try:
            with torch.device(device):
                new = build()
            break
        except Exception as e:  # try next signature
            errors.append(e)
    if new is None:
        raise RuntimeError(f"could not rebuild {cls.__name__} from config: {errors[-1]}")
    new = new.to(device)
    src = module.state_dict()
    dst = new.state_dict()
    missing = [k for k in dst if k not in src]
    if missing:
        raise RuntimeError(f"rebuilt {cls.__name__} misses {len(missing)} keys (e.g. {missing[0]})")
    with torch.no_grad():
        for k, v in dst.items():
            s = src[k]
            if v.shape != s.shape:
                raise RuntimeError(f"shape mismatch for {k}: {tuple(v.shape)} vs {tuple(s.shape)}")
            if v.dtype != dtypes.storage_dtype_for(s.dtype, device):
                # keep the source's storage dtype (e.g. fp8 weights stay fp8 on B200)
                _assign_by_key(new, k, _move_tensor(s, device))
            else:
                v.copy_(s)
    return new


def _assign_by_key(root: nn.Module, key: str, value: torch.Tensor) -> None:
    *path, leaf = key.split(".")
    mod = root
    for p in path:
        mod = getattr(mod, p)
    if leaf in mod._parameters:
        mod._parameters[leaf] = nn.Parameter(value, requires_grad=False)
    else:
        mod._buffers[leaf] = value


def _finalize_replica(replica: nn.Module, device: torch.device, safe_attention: bool) -> nn.Module:
    memory.clear_model_caches(replica, quiet=True)
    if not dtypes.device_supports_float8(device):
        with torch.no_grad():
            for m in replica.modules():
                for name, p in list(m._parameters.items()):
                    if p is not None and dtypes.is_float8_dtype(p.dtype):
                        m._parameters[name] = nn.Parameter(p.detach().to(torch.float16), requires_grad=False)
                for name, b in list(m._buffers.items()):
                    if b is not None and dtypes.is_float8_dtype(b.dtype):
                        m._buffers[name] = b.to(torch.float16)
    for m in replica.modules():
        if hasattr(m, "gradient_checkpointing"):
            try:
                m.gradient_checkpointing = False
            except Exception:
                pass
        if getattr(m, "_gradient_checkpointing_func", None) is not None:
            try:
                m._gradient_checkpointing_func = None
            except Exception:
                pass
        if hasattr(m, "_hf_hook"):
            try:
                delattr(m, "_hf_hook")
            except Exception:
                pass
        # plain hook dicts copied by deepcopy would keep references to the source
        for hk in ("_forward_hooks", "_forward_pre_hooks", "_backward_hooks"):
            d = getattr(m, hk, None)
            if d:
                d.clear()
    replica.eval()
    for p in replica.parameters():
        p.requires_grad_(False)
    for m in replica.modules():  # stray buffers
        for name, b in list(m._buffers.items()):
            if b is not None and b.device != device:
                m._buffers[name] = b.to(device)
    if safe_attention:
        memory.disable_flash_xformers(replica)
    return replica


def safe_model_clone(source: nn.Module, device: DeviceLike, safe_attention: bool = False,
                     index: Optional[int] = None) -> nn.Module:
    """Produce an inference-only replica of ``source`` on ``device``.

    Same device -> the source object itself (ADP:594-597).  Otherwise try, in order:
    D2D deepcopy, rebuild-from-config, structural clone.  CUDA OOM is re-raised so the
    caller can skip the device (ADP:1114-1120)."""
    device = torch.device(device)
    faults.check_setup(str(device), index)
    src_dev = memory.module_device(source)
    if src_dev is not None and src_dev == device:
        return source
    method = "d2d-deepcopy"
    try:
        replica = clone_module_d2d(source, device)
    except torch.cuda.OutOfMemoryError:
        raise
    except Exception as e1:
        if "out of memory" in str(e1).lower():
            raise
        log.debug("d2d clone failed (%s); trying rebuild-from-config", e1)
        try:
            method = "rebuild-from-config"
            replica = clone_module_from_config(source, device)
        except torch.cuda.OutOfMemoryError:
            raise
        except Exception as e2:
            if "out of memory" in str(e2).lower():
                raise
            log.debug("rebuild failed (%s); trying structural clone", e2)
            method = "structural"
            replica = clone_module_structural(source, device)
    replica = _finalize_replica(replica, device, safe_attention)
    log.info("Cloned %s to %s via %s (%.1f MiB)", type(source).__name__, device, method,
             memory.module_bytes(replica) / 2 ** 20)
    return replica
