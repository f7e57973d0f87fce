"""Replication helpers."""
import dataclasses

import torch
import torch.nn as nn

from comfyui_parallelanything_b200.models import flux, unet
from comfyui_parallelanything_b200.utils import dtypes, memory, replicate


def test_fp8_predicates():
    assert dtypes.is_float8_dtype(torch.float8_e4m3fn) and dtypes.is_float8_dtype(torch.float8_e5m2)
    assert not dtypes.is_float8_dtype(torch.bfloat16) and not dtypes.is_float8_dtype(None)
    assert dtypes.check_sm80_support("cpu") and not dtypes.device_supports_float8("cpu")
    assert dtypes.storage_dtype_for(torch.float8_e4m3fn, "cpu") == torch.float16


def test_extract_config_sources():
    m = flux.Flux(flux.flux_tiny_params())
    cfg = replicate.extract_model_config(m)
    assert cfg["hidden_size"] == 256 and cfg["depth"] == 2 and cfg["axes_dim"] == [16, 56, 56]
    u = unet.UNetModel(**unet.tiny_config())
    ucfg = replicate.extract_model_config(u)
    assert ucfg["model_channels"] == 32 and ucfg["channel_mult"] == [1, 2]


def test_clone_dataclass():
    @dataclasses.dataclass
    class P:
        a: int
        t: torch.Tensor
        l: list
    p = P(1, torch.ones(2), [torch.zeros(1), 3])
    q = replicate.clone_dataclass_or_object(p)
    assert q is not p and q.t is not p.t and torch.equal(q.t, p.t) and q.l[1] == 3


def _same(a: nn.Module, b: nn.Module):
    sa, sb = a.state_dict(), b.state_dict()
    assert sa.keys() == sb.keys()
    for k in sa:
        assert torch.equal(sa[k].float(), sb[k].float()), k


def test_three_clone_strategies_agree():
    m = flux.Flux(flux.flux_tiny_params()).eval()
    m.img_ids = torch.zeros(3)                      # a device-bound cache that must not follow
    for fn in (replicate.clone_module_d2d, replicate.clone_module_from_config, replicate.clone_module_structural):
        c = fn(m, "cpu")
        assert c is not m
        _same(m, c)
        assert next(c.parameters()).data_ptr() != next(m.parameters()).data_ptr()
    c = replicate._finalize_replica(replicate.clone_module_d2d(m, "cpu"), torch.device("cpu"), False)
    assert c.img_ids is None and not any(p.requires_grad for p in c.parameters()) and not c.training


def test_shared_parameters_stay_shared():
    class Tied(nn.Module):
        def __init__(self):
            super().__init__()
            self.a = nn.Linear(3, 3)
            self.b = nn.Linear(3, 3)
            self.b.weight = self.a.weight
        def forward(self, x):
            return self.b(self.a(x))
    c = replicate.clone_module_d2d(Tied(), "cpu")
    assert c.a.weight is c.b.weight


def test_fp8_downcast_on_non_fp8_device():
    m = nn.Linear(4, 4)
    m.weight = nn.Parameter(m.weight.detach().to(torch.float8_e4m3fn), requires_grad=False)
    c = replicate.clone_module_d2d(m, "cpu")
    assert c.weight.dtype == torch.float16          # cpu has no fp8 support -> widened (ADP:403-404)


def test_safe_clone_same_device_returns_source():
    m = nn.Linear(2, 2)
    assert replicate.safe_model_clone(m, "cpu") is m


def test_clear_caches_and_disable_flash():
    m = flux.Flux(flux.flux_tiny_params())
    m.freqs_cis = torch.ones(1)
    m.double_blocks[0].kv_cache = torch.ones(1)
    assert memory.clear_model_caches(m, quiet=True) == 2 and m.freqs_cis is None
    m.double_blocks[0].img_attn.use_flash_attention = True
    m.use_xformers = True
    assert memory.disable_flash_xformers(m) >= 2
    assert m.double_blocks[0].img_attn.use_flash_attention is False and m.use_xformers is False
    assert memory.get_free_vram("cpu") == 0


def test_ingest_state_dict_with_prefix(tmp_path):
    from comfyui_parallelanything_b200.exec import pack_cache
    src = flux.Flux(flux.flux_tiny_params())
    sd = {"model.diffusion_model." + k: v for k, v in src.state_dict().items()}
    path = tmp_path / "ckpt.pt"
    torch.save(sd, path)
    dst = flux.Flux(flux.flux_tiny_params())
    res = pack_cache.ingest_state_dict(dst, str(path), prefix="model.diffusion_model.", strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    _same(src, dst)


def test_packed_weight_checkpoint_roundtrip_and_mismatch(tmp_path):
    """SURVEY §5 checkpoint / resume of the executors' packed weight tables (exec/pack_cache.py), on a stand-in
    executor so it runs without a GPU."""
    import pytest
    import torch
    from comfyui_parallelanything_b200.exec import pack_cache

    class FakeExec:
        fp8 = False

        def __init__(self, seed):
            g = torch.Generator().manual_seed(seed)
            self.W = {"a.w": torch.randn(4, 8, generator=g).bfloat16(), "a.b": None,
                      "q.bytes": torch.randint(0, 255, (16,), generator=g, dtype=torch.uint8), "tile": 224}

    src, dst = FakeExec(1), FakeExec(2)
    path = str(tmp_path / "packed.pa")
    assert pack_cache.save_packed(src, path) > 0
    meta = pack_cache.load_packed_into(dst, path)
    assert meta["class"] == "FakeExec" and meta["keys"] == 4
    assert torch.equal(dst.W["a.w"], src.W["a.w"]) and torch.equal(dst.W["q.bytes"], src.W["q.bytes"])
    assert dst.W["a.b"] is None and dst.W["tile"] == 224

    class OtherExec(FakeExec):
        pass
    with pytest.raises(ValueError):
        pack_cache.load_packed_into(OtherExec(3), path)                  # wrong executor class
    bad = FakeExec(4)
    bad.W["extra"] = torch.zeros(1)
    with pytest.raises(KeyError):
        pack_cache.load_packed_into(bad, path)                           # key set mismatch
    shp = FakeExec(5)
    shp.W["a.w"] = torch.zeros(2, 2).bfloat16()
    with pytest.raises(ValueError):
        pack_cache.load_packed_into(shp, path)                           # shape mismatch
    with pytest.raises(TypeError):
        pack_cache.save_packed(object(), path)                           # no packed table


def test_executor_shell_clone_structure_and_p2p_fill():
    """parallel/replicate_nvl.shell_like: the receive side of a weight replication - same packed-table keys / shapes on
    the target device, fresh runtime state, aliasing between the flat block list and the layer tree preserved."""
    import torch
    import torch.nn as nn
    from types import SimpleNamespace
    from comfyui_parallelanything_b200.exec.graphs import GraphCache
    from comfyui_parallelanything_b200.exec.pack_cache import packed_table
    from comfyui_parallelanything_b200.parallel import replicate_nvl

    class Blk:
        def __init__(self, i):
            self.w = torch.full((4, 4), float(i))
            self.heads = 2
            self._kv = torch.zeros(3)
            self._kv_sig = ("x",)

    class FakeExec(nn.Module):
        pa_native = True

        def __init__(self):
            super().__init__()
            self.device = torch.device("cpu")
            self.params = SimpleNamespace(dim=8, axes=[1, 2])
            self.inp = [[("st", Blk(1))], [("st", Blk(2))]]
            self._tblocks = [self.inp[0][0][1], self.inp[1][0][1]]
            self.emb_w = torch.arange(6.0)
            self._graphs = GraphCache("cpu", enabled=False)
            self._io = {"k": 1}
            self.launches_per_step = 7

    src = FakeExec()
    sh = replicate_nvl.shell_like(src, "cpu")
    assert type(sh) is FakeExec and sh is not src and sh.launches_per_step == 7 and sh._io == {}
    assert sh.params.dim == 8 and sh.params is not src.params and sh._graphs is not src._graphs
    assert sh._tblocks[0] is sh.inp[0][0][1] and sh._tblocks[1] is sh.inp[1][0][1]          # aliasing preserved
    assert sh._tblocks[0]._kv is None and sh._tblocks[0]._kv_sig is None and sh._tblocks[0].heads == 2
    ts, td = packed_table(src), packed_table(sh)
    assert list(ts) == list(td) and all(td[k].shape == ts[k].shape and td[k] is not ts[k] for k in ts)
    pairs = replicate_nvl._pairs(src, [sh])
    for s_, ds in pairs:
        ds[0].copy_(s_)
    assert all(torch.equal(td[k], ts[k]) for k in ts)
    # chunk packing: every byte of every tensor exactly once, offsets aligned, slots never overflow
    big = [(torch.empty(1000, dtype=torch.uint8), []), (torch.empty(70000, dtype=torch.uint8), []),
           (torch.empty(16, dtype=torch.uint8), [])]
    seen = {0: 0, 1: 0, 2: 0}
    for chunk in replicate_nvl._chunks(big, 32768):
        for idx, pos, take, off in chunk:
            assert off % 256 == 0 and off + take <= 32768 and pos == seen[idx]
            seen[idx] += take
    assert seen == {0: 1000, 1: 70000, 2: 16}
