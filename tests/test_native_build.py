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


def test_glu_weight_interleaving_matches_epilogue_contract():
    """geglu / swiglu GEMM epilogues read output column n/2 from rows [a(32) | g(32)] of each 64-row group."""
    import torch
    from comfyui_parallelanything_b200 import ops
    wa = torch.arange(64 * 4, dtype=torch.float32).view(64, 4)
    wg = -wa
    w = ops.interleave_glu(wa, wg)
    assert w.shape == (128, 4)
    assert torch.equal(w[0:32], wa[0:32]) and torch.equal(w[32:64], wg[0:32])
    assert torch.equal(w[64:96], wa[32:64]) and torch.equal(w[96:128], wg[32:64])
    b = ops.interleave_glu(torch.arange(64.0), -torch.arange(64.0))
    assert torch.equal(b[:32], torch.arange(32.0)) and torch.equal(b[32:64], -torch.arange(32.0))


def test_zimage_family_is_registered_for_the_native_executor():
    import torch
    from comfyui_parallelanything_b200 import exec as pa_exec
    from comfyui_parallelanything_b200.models import zimage
    m = zimage.ZImageModel(zimage.zimage_tiny_params())
    assert pa_exec.builder_for(m) is not None
    ids = zimage.ZImageModel.make_ids(2, 3, 2, 2, "cpu")
    assert ids.shape == (2, 7, 3)
    assert ids[0, :3, 0].tolist() == [1.0, 2.0, 3.0] and ids[0, 3:, 0].unique().tolist() == [4.0]
    assert ids[0, 3:, 1].tolist() == [0.0, 0.0, 1.0, 1.0] and ids[0, 3:, 2].tolist() == [0.0, 1.0, 0.0, 1.0]
    # head dim 64: no native schedule -> the engine falls back to a torch replica
    assert pa_exec.builder_for(zimage.ZImageModel(zimage.ZImageParams(dim=256, n_heads=4, n_layers=1, n_refiner_layers=1,
                                                                      ffn_hidden=256, cap_feat_dim=64, adaln_dim=64,
                                                                      axes_dims=[16, 24, 24]))) is None


def test_sass_uses_blackwell_paths_and_has_no_issue_waterfalls():
    """Compile-time regression guard (needs only cuobjdump): the built library must contain the sm_100a tensor-core /
    TMA instructions the design rests on - including the CTA-pair forms - and none of the ELECT / R2UR.BROADCAST /
    BRA.U.ANY waterfall loops ptxas wraps around tcgen05 / TMA issue inside a ``lane == 0`` branch (~90 cycles per MMA,
    DESIGN.md §5 "issue path")."""
    import os
    import shutil
    import subprocess

    import pytest
    so = os.path.join(os.path.dirname(ops.__file__), "_C.so")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(so) or not os.path.exists(cuobjdump):
        pytest.skip("no built library / cuobjdump")
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "UTMALDG", "UTMALDG.3D.2CTA", "UTCBAR.2CTA.MULTICAST",
                     "LDTM", "STTM", "UTCCP"):
        assert mnemonic in sass, f"{mnemonic} missing from the SASS"
    assert "BRA.U.ANY" not in sass, "tcgen05 / TMA issue fell back to a per-instruction waterfall loop"


def test_executor_workspace_cache_is_lru_bounded():
    from comfyui_parallelanything_b200 import exec as pa_exec
    cache, evicted = {}, []
    for i in range(pa_exec.WORKSPACE_LIMIT + 3):
        pa_exec.cache_workspace(cache, ("shape", i), {"i": i}, on_evict=evicted.append)
    assert len(cache) == pa_exec.WORKSPACE_LIMIT and evicted == [("shape", 0), ("shape", 1), ("shape", 2)]
    assert pa_exec.touch_workspace(cache, ("shape", 3))["i"] == 3            # a hit makes it the most recent entry
    pa_exec.cache_workspace(cache, ("shape", 99), {"i": 99}, on_evict=evicted.append)
    assert ("shape", 3) in cache and evicted[-1] == ("shape", 4)
    assert pa_exec.touch_workspace(cache, ("missing",)) is None
