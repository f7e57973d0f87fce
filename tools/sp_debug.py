"""Sequence-parallel debugging aid (2 GPUs): per step and GPU the epoch counter, the watchdog error word and the first flag
slots, plus the error vs the fp32 oracle.   G=0 disables graphs, STEPS=n runs n steps."""
import copy, os, sys, torch  # noqa: E401
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comfyui_parallelanything_b200 as pa
from comfyui_parallelanything_b200.models import flux
from comfyui_parallelanything_b200.utils.config import EngineConfig
devs = ["cuda:0", "cuda:1"]
torch.manual_seed(0)
p = flux.FluxParams(in_channels=64, out_channels=64, vec_in_dim=768, context_in_dim=512, hidden_size=512,
                    mlp_ratio=4.0, num_heads=4, depth=2, depth_single_blocks=2)
m = flux.Flux(p).to(device=devs[0], dtype=torch.bfloat16).eval()
oracle = copy.deepcopy(m).float()
chain = None
for d in devs:
    chain = pa.ParallelDevice().add_device(d, 50.0, chain)[0]
cfg = EngineConfig(batch1_mode="ulysses", flag_timeout_ms=3000, cuda_graphs=os.environ.get("G", "1") == "1")
pa.ParallelAnything().setup_parallel(m, chain, config=cfg)
eng = m._parallel_engine
sp = eng._ulysses
inp = flux.example_inputs(p, 1, 256, 256, txt_len=64, device=devs[0], dtype=torch.bfloat16)
for it in range(int(os.environ.get("STEPS", "1"))):
    with torch.no_grad():
        got = m(inp["x"], inp["timesteps"], context=inp["context"], y=inp["y"], guidance=inp["guidance"])
    for d in devs:
        torch.cuda.synchronize(d)
    for g in range(2):
        print("step", it, "gpu", g, "epoch", sp.epoch[g].item(), "err", hex(sp.err[g].item() & 0xffffffff), "flags", sp.flags[g][:24].tolist())
    want = oracle(**{k: v.float() for k, v in inp.items()})
    print("rel", ((got.float() - want).abs().mean() / want.abs().mean()).item())
