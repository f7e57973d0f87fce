"""comfyui-parallelanything_b200 — a Blackwell-native multi-GPU diffusion inference engine
with the node API of FearL0rd/ComfyUI-ParallelAnything.

ComfyUI loads this directory as a custom node and reads the two mappings below
(reference: /root/reference/__init__.py:1-3, any_device_parallel.py:1473-1483).
"""
from .nodes import (NODE_CLASS_MAPPINGS, NODE_DISPLAY_NAME_MAPPINGS, ParallelAnything, ParallelDevice,
                    ParallelDeviceList)
from .engine import ParallelEngine, cleanup_parallel_model
from .utils.config import EngineConfig

__version__ = "0.1.0"
__all__ = [
    "NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS", "ParallelAnything", "ParallelDevice",
    "ParallelDeviceList", "ParallelEngine", "EngineConfig", "cleanup_parallel_model",
]
