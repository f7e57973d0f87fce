"""ComfyUI custom-node entry point when the repository root itself is dropped into ``custom_nodes/``
(reference: /root/reference/__init__.py:1-3)."""
try:
    from .comfyui_parallelanything_b200 import NODE_CLASS_MAPPINGS, NODE_DISPLAY_NAME_MAPPINGS
except ImportError:  # imported as a top-level module (pytest rootdir, scripts)
    from comfyui_parallelanything_b200 import NODE_CLASS_MAPPINGS, NODE_DISPLAY_NAME_MAPPINGS

__all__ = ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS"]
