"""The ComfyUI-facing path on real GPUs: ParallelAnything.setup_parallel -> native sm_100a executors behind
the hooked forward, single GPU and (when available) multi-GPU with in-kernel NVLink scatter/gather."""
import copy

import pytest
import torch
import torch.nn as nn

import comfyui_parallelanything_b200 as pa
from comfyui_parallelanything_b200.models import flux, unet

pytestmark = pytest.mark.gpu


def _chain(devs, pcts=None):
    c = None
    for i, d in enumerate(devs):
        c = pa.ParallelDevice().add_device(d, (pcts or [100.0 / len(devs)] * len(devs))[i], c)[0]
    return c


def _flux_case(devs, pcts=None, batch=4):
    torch.manual_seed(0)
    p = flux.FluxParams(in_channels=64, out_channels=64, vec_in_dim=768, context_in_dim=512, hidden_size=512,
                        mlp_ratio=4.0, num_heads=4, depth=1, depth_single_blocks=2)
    m = flux.Flux(p).to(device=devs[0], dtype=torch.bfloat16).eval()
    oracle = copy.deepcopy(m).float()
    out, = pa.ParallelAnything().setup_parallel(m, _chain(devs, pcts))
    assert out is m and m._true_parallel_active
    eng = m._parallel_engine
    assert all(getattr(r, "pa_native", False) for r in eng.replicas.values()), "expected native executors on B200"
    inp = flux.example_inputs(p, batch, 256, 256, txt_len=64, device=devs[0], dtype=torch.bfloat16)
    with torch.no_grad():
        for _ in range(4):          # calls 3+ replay the CUDA graph captured on every replica's own device;
            inp["x"].mul_(0.97)     # the latent changes in place, so a replay that did nothing would be detected
            got = m(inp["x"], inp["timesteps"], context=inp["context"], y=inp["y"], guidance=inp["guidance"])
            torch.cuda.synchronize()
            got = got.clone()
        want = oracle(**{k: v.float() for k, v in inp.items()})
    torch.cuda.synchronize()
    rel = (got.float() - want).abs().mean().item() / want.abs().mean().item()
    pa.cleanup_parallel_model(m)
    return rel, eng


def test_single_gpu_native_flux():
    rel, _ = _flux_case(["cuda:0"])
    assert rel < 0.03, rel


def test_native_unet_through_nodes():
    cfg = unet.mini_sdxl_config()
    torch.manual_seed(1)
    m = unet.UNetModel(**cfg).to(device="cuda:0", dtype=torch.bfloat16).eval()
    oracle = copy.deepcopy(m).float()
    pa.ParallelAnything().setup_parallel(m, _chain(["cuda:0"]))
    inp = unet.example_inputs(cfg, 2, 256, 256, ctx_len=77, device="cuda:0", dtype=torch.bfloat16)
    with torch.no_grad():
        got = m(inp["x"], inp["timesteps"], context=inp["context"], y=inp["y"])
        want = oracle(**{k: v.float() for k, v in inp.items()})
    rel = (got.float() - want).abs().mean().item() / want.abs().mean().item()
    pa.cleanup_parallel_model(m)
    assert rel < 0.04, rel


@pytest.mark.multigpu
def test_multi_gpu_fused_dp_matches_single():
    n = min(torch.cuda.device_count(), 4)
    devs = [f"cuda:{i}" for i in range(n)]
    rel, eng = _flux_case(devs, batch=2 * n + 1)
    assert rel < 0.03, rel
    assert any(r.get("fused") for r in eng.metrics.rows), "fused in-process path was not taken"
    # replicas 1.. were filled over NVLink from the lead's packed weights (multicast kernel or peer copies), not re-packed
    rep = eng.setup_report.get("replication")
    assert rep and rep["method"] in ("nvls", "p2p") and rep["receivers"] == n - 1, rep
    assert eng.metrics.counters.get("native_graph_steps", 0) >= 1, "graphs were not replayed by the native host threads"


@pytest.mark.multigpu
def test_fresh_tensors_every_step_still_replay_graphs():
    """A sampler hands the hooked forward NEW tensors every step (and ComfyUI re-concatenates the conditioning): the
    engine stages them into fixed buffers, so every replica keeps replaying ONE captured graph."""
    devs = [f"cuda:{i}" for i in range(2)]
    torch.manual_seed(0)
    p = flux.FluxParams(in_channels=64, out_channels=64, vec_in_dim=768, context_in_dim=512, hidden_size=512,
                        mlp_ratio=4.0, num_heads=4, depth=1, depth_single_blocks=2)
    m = flux.Flux(p).to(device=devs[0], dtype=torch.bfloat16).eval()
    oracle = copy.deepcopy(m).float()
    pa.ParallelAnything().setup_parallel(m, _chain(devs))
    eng = m._parallel_engine
    base = flux.example_inputs(p, 4, 256, 256, txt_len=64, device=devs[0], dtype=torch.bfloat16)
    outs = []
    with torch.no_grad():
        for it in range(6):
            inp = {k: (v * (1.0 - 0.05 * it)).clone() if k == "x" else v.clone() for k, v in base.items()}   # all fresh
            got = m(inp["x"], inp["timesteps"], context=inp["context"], y=inp["y"], guidance=inp["guidance"])
            outs.append((inp, got))
        torch.cuda.synchronize()
        for inp, got in outs[-2:]:
            want = oracle(**{k: v.float() for k, v in inp.items()})
            rel = (got.float() - want).abs().mean().item() / want.abs().mean().item()
            assert rel < 0.03, rel
    assert not torch.equal(outs[-1][1], outs[-2][1])              # results were not overwritten by later steps
    graphs = {n_: len(r._graphs) for n_, r in eng.replicas.items()}
    assert all(v == 1 for v in graphs.values()), graphs
    assert eng.metrics.counters.get("native_graph_steps", 0) >= 3
    pa.cleanup_parallel_model(m)


@pytest.mark.multigpu
def test_multi_gpu_weighted_split_and_generic_module():
    class Toy(nn.Module):
        def __init__(self):
            super().__init__()
            self.l = nn.Linear(16, 16)

        def forward(self, x, timesteps, context=None, **kw):
            return torch.tanh(self.l(x)) + timesteps[:, None] + context.mean(1)
    m = Toy().to("cuda:0").eval()
    plain = copy.deepcopy(m)
    pa.ParallelAnything().setup_parallel(m, _chain(["cuda:0", "cuda:1"], [70, 30]))
    x, t, c = torch.randn(10, 16, device="cuda:0"), torch.rand(10, device="cuda:0"), torch.randn(10, 3, 16, device="cuda:0")
    with torch.no_grad():
        got, want = m(x, t, context=c), plain(x, t, context=c)
    torch.cuda.synchronize()
    assert torch.allclose(got, want, atol=1e-5)
    assert m._parallel_engine.metrics.rows[-1]["sizes"] == [7, 3]
    pa.cleanup_parallel_model(m)


def test_stranded_model_is_moved_back_to_its_load_device():
    """C21 (ADP:932-961): weights left on the CPU by an earlier run while the patcher says cuda:0."""
    from comfyui_parallelanything_b200 import nodes

    class Patcher:
        load_device = torch.device("cuda:0")

    m = nn.Sequential(nn.Linear(8, 8), nn.Linear(8, 8))
    nodes._repair_stranded(Patcher(), m)
    assert all(p.device == torch.device("cuda:0") for p in m.parameters())


def _zimage_case(devs, batch):
    from comfyui_parallelanything_b200.models import zimage
    torch.manual_seed(2)
    p = zimage.zimage_tiny_params()
    m = zimage.ZImageModel(p).to(device=devs[0], dtype=torch.bfloat16).eval()
    oracle = copy.deepcopy(m).float()
    pa.ParallelAnything().setup_parallel(m, _chain(devs))
    eng = m._parallel_engine
    assert all(getattr(r, "pa_native", False) for r in eng.replicas.values()), "expected native Z-Image executors"
    inp = zimage.example_inputs(p, batch, 256, 256, cap_len=32, device=devs[0], dtype=torch.bfloat16)
    with torch.no_grad():
        for _ in range(2):                  # second call: cached caption path
            got = m(inp["x"], inp["timesteps"], context=inp["context"])
        want = oracle(**{k: v.float() for k, v in inp.items()})
    torch.cuda.synchronize()
    rel = (got.float() - want).abs().mean().item() / want.abs().mean().item()
    pa.cleanup_parallel_model(m)
    return rel, eng


def test_native_zimage_through_nodes():
    rel, _ = _zimage_case(["cuda:0"], 3)
    assert rel < 0.03, rel


@pytest.mark.multigpu
def test_multi_gpu_zimage_fused():
    n = min(torch.cuda.device_count(), 4)
    rel, eng = _zimage_case([f"cuda:{i}" for i in range(n)], 2 * n + 1)
    assert rel < 0.03, rel
    assert any(r.get("fused") for r in eng.metrics.rows), "fused in-process path was not taken"


@pytest.mark.multigpu
@pytest.mark.parametrize("fp8", [False, True])
def test_batch1_sequence_parallel_ulysses(fp8, monkeypatch):
    """batch == 1: every GPU of the chain works on the one sample (token-sliced linears, head-sliced attention, peer-pull
    all-to-all) instead of the reference's sequential layer split; result vs the fp32 oracle, graphs replayed."""
    from comfyui_parallelanything_b200.utils.config import EngineConfig
    n = 2
    devs = [f"cuda:{i}" for i in range(n)]
    torch.manual_seed(0)
    p = flux.FluxParams(in_channels=64, out_channels=64, vec_in_dim=768, context_in_dim=512, hidden_size=512,
                        mlp_ratio=4.0, num_heads=4, depth=2, depth_single_blocks=2)
    m = flux.Flux(p).to(device=devs[0], dtype=torch.bfloat16).eval()
    oracle = copy.deepcopy(m).float()
    cfg = EngineConfig(fp8=fp8, batch1_mode="ulysses")
    pa.ParallelAnything().setup_parallel(m, _chain(devs), config=cfg)
    eng = m._parallel_engine
    assert eng._ulysses is not None, "sequence-parallel path was not set up"
    base = flux.example_inputs(p, 1, 256, 256, txt_len=64, device=devs[0], dtype=torch.bfloat16)
    rels = []
    with torch.no_grad():
        for it in range(5):
            inp = {k: (v * (1.0 - 0.1 * it)).clone() if k == "x" else v.clone() for k, v in base.items()}
            got = m(inp["x"], inp["timesteps"], context=inp["context"], y=inp["y"], guidance=inp["guidance"])
            want = oracle(**{k: v.float() for k, v in inp.items()})
            torch.cuda.synchronize()
            rels.append((got.float() - want).abs().mean().item() / want.abs().mean().item())
    eng._ulysses.check_error()
    assert max(rels) < (0.05 if fp8 else 0.03), rels
    assert eng.metrics.counters.get("ulysses_steps", 0) == 5
    assert all(len(s.replica._graphs) >= 1 for s in eng.slots)
    pa.cleanup_parallel_model(m)


@pytest.mark.multigpu
@pytest.mark.parametrize("fp8", [False, True])
def test_batch1_sequence_parallel_ulysses_wan(fp8):
    """The video family at batch 1 (its usual workload): token-sliced linears + q/k norm + RoPE, head-sliced
    self-attention between two peer-pull exchanges, local text cross-attention; result vs the fp32 oracle, a new latent
    tensor every step, graphs replayed from the third step on."""
    from comfyui_parallelanything_b200.models import wan
    from comfyui_parallelanything_b200.utils.config import EngineConfig
    n = 2
    devs = [f"cuda:{i}" for i in range(n)]
    torch.manual_seed(0)
    p = wan.wan_tiny_params()
    m = wan.WanModel(p).to(device=devs[0], dtype=torch.bfloat16).eval()
    oracle = copy.deepcopy(m).float()
    cfg = EngineConfig(fp8=fp8, batch1_mode="ulysses")
    pa.ParallelAnything().setup_parallel(m, _chain(devs), config=cfg)
    eng = m._parallel_engine
    assert eng._ulysses is not None and eng._ulysses.family == "wan", "sequence-parallel WAN path was not set up"
    base = wan.example_inputs(p, 1, frames=8, height=64, width=64, device=devs[0], dtype=torch.bfloat16)
    rels = []
    with torch.no_grad():
        for it in range(5):
            inp = {k: (v * (1.0 - 0.1 * it)).clone() if k == "x" else v.clone() for k, v in base.items()}
            got = m(inp["x"], inp["timesteps"], context=inp["context"])
            want = oracle(**{k: v.float() for k, v in inp.items()})
            torch.cuda.synchronize()
            rels.append((got.float() - want).abs().mean().item() / want.abs().mean().item())
    eng._ulysses.check_error()
    assert max(rels) < (0.05 if fp8 else 0.03), rels
    assert eng.metrics.counters.get("ulysses_steps", 0) == 5
    assert all(len(s.replica._graphs) >= 1 for s in eng.slots)
    pa.cleanup_parallel_model(m)


@pytest.mark.multigpu
@pytest.mark.parametrize("fp8", [False, True])
def test_batch1_sequence_parallel_ulysses_zimage(fp8):
    """NextDiT at batch 1: image-only exchange tables for the noise refiner, joint [caption | image] tables for the main
    layers, caption path replicated and cached; result vs the fp32 oracle with a new latent every step."""
    from comfyui_parallelanything_b200.models import zimage
    from comfyui_parallelanything_b200.utils.config import EngineConfig
    n = 2
    devs = [f"cuda:{i}" for i in range(n)]
    torch.manual_seed(0)
    p = zimage.zimage_tiny_params()
    m = zimage.ZImageModel(p).to(device=devs[0], dtype=torch.bfloat16).eval()
    oracle = copy.deepcopy(m).float()
    cfg = EngineConfig(fp8=fp8, batch1_mode="ulysses")
    pa.ParallelAnything().setup_parallel(m, _chain(devs), config=cfg)
    eng = m._parallel_engine
    assert eng._ulysses is not None and eng._ulysses.family == "zimage", "sequence-parallel Z-Image path was not set up"
    base = zimage.example_inputs(p, 1, 256, 256, cap_len=32, device=devs[0], dtype=torch.bfloat16)
    rels = []
    with torch.no_grad():
        for it in range(5):
            inp = {k: (v * (1.0 - 0.1 * it)).clone() if k == "x" else v.clone() for k, v in base.items()}
            got = m(inp["x"], inp["timesteps"], context=inp["context"])
            want = oracle(**{k: v.float() for k, v in inp.items()})
            torch.cuda.synchronize()
            rels.append((got.float() - want).abs().mean().item() / want.abs().mean().item())
    eng._ulysses.check_error()
    assert max(rels) < (0.06 if fp8 else 0.03), rels
    assert eng.metrics.counters.get("ulysses_steps", 0) == 5
    assert all(len(s.replica._graphs) >= 1 for s in eng.slots)
    pa.cleanup_parallel_model(m)
