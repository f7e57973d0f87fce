"""The native library must at least build and import on the GPU-less dev box."""
import torch

from comfyui_parallelanything_b200 import ops


def test_extension_imports_on_cpu():
    assert ops.available(), f"_C failed to import: {ops.load_error()!r} (run python tools/build.py)"
    C = ops.require()
    for name in ("gemm", "attention", "layernorm_modulate", "groupnorm_silu", "cfg_euler_store", "signal_flags",
                 "wait_flags", "ipc_get_handle", "HostExecutor"):
        assert hasattr(C, name), name
    assert ops.EPI["qkv_rope"] == C.EPI_QKV_ROPE and ops.EPI["euler_unpatch"] == C.EPI_EULER_UNPATCH
    assert not ops.native_ok("cpu")
    if not torch.cuda.is_available():
        assert not ops.native_ok("cuda:0")
