"""Stage-by-stage comparison of the WAN executor against the fp32 oracle (block 0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from comfyui_parallelanything_b200.exec.wan_exec import WanExecutor
from comfyui_parallelanything_b200.models import wan, flux

dev = torch.device("cuda:0")
p = wan.wan_tiny_params()
torch.manual_seed(2)
m = wan.WanModel(p).to(device=dev, dtype=torch.bfloat16).eval()
ex = WanExecutor(m, dev)
o = wan.WanModel(p).to(device=dev, dtype=torch.float32).eval()
o.load_state_dict({k: v.float() for k, v in m.state_dict().items()})
inp = wan.example_inputs(p, 2, frames=8, height=128, width=192, device=dev, dtype=torch.bfloat16)
ex._dbg = {}
with torch.no_grad():
    got = ex(**inp)
d = ex._dbg


def rel(a, b, name):
    a, b = a.float(), b.float()
    print(f"{name:10s} mean_rel {((a-b).abs().mean()/(b.abs().mean()+1e-9)).item():.5f}  max_abs {(a-b).abs().max().item():.4f}  ref_mean {b.abs().mean().item():.4f}")


with torch.no_grad():
    x = inp["x"].float(); t = inp["timesteps"].float(); ctx_in = inp["context"].float()
    xe = o.patch_embedding(x); grid = xe.shape[2:]; xt = xe.flatten(2).transpose(1, 2)
    rel(d["x0"], xt, "patch")
    e = o.time_embedding(flux.timestep_embedding(t, 256, time_factor=1.0))
    rel(d["e"], e, "e")
    e0 = o.time_projection(e)
    rel(d["e0"], e0, "e0")
    ctx = o.text_embedding(ctx_in)
    rel(d["ctx"], ctx, "ctx")
    blk = o.blocks[0]
    ee = (blk.modulation + e0.unflatten(1, (6, p.dim))).chunk(6, dim=1)
    xm = blk.norm1(xt) * (1 + ee[1]) + ee[0]
    rel(d["xm0"], xm, "xm0")
    sa = blk.self_attn
    q, k, v = sa.q(xm), sa.k(xm), sa.v(xm)
    rel(d["qkv_raw"], torch.cat([q, k, v], -1), "qkv_raw")
    freqs = o.rope_embedder(o.make_ids(2, *grid, dev))
    b, s_, n, dd = 2, xm.shape[1], p.num_heads, 128
    qn = sa.norm_q(q).view(b, s_, n, dd).transpose(1, 2); kn = sa.norm_k(k).view(b, s_, n, dd).transpose(1, 2)
    qr, kr = flux.apply_rope(qn, kn, freqs)
    ref = torch.cat([qr.transpose(1, 2).reshape(b, s_, -1), kr.transpose(1, 2).reshape(b, s_, -1), v], -1)
    rel(d["qkv_rope"][..., :p.dim], ref[..., :p.dim], "q_rope")
    rel(d["qkv_rope"][..., p.dim:2*p.dim], ref[..., p.dim:2*p.dim], "k_rope")
    att = torch.nn.functional.scaled_dot_product_attention(qr, kr, v.view(b, s_, n, dd).transpose(1, 2)).transpose(1, 2).reshape(b, s_, -1)
    rel(d["att"], att, "att")
    x1 = xt + sa.o(att) * ee[2]
    rel(d["x_sa"], x1, "x_sa")
    x2 = x1 + blk.cross_attn(blk.norm3(x1), ctx)
    rel(d["x_ca"], x2, "x_ca")
    y = blk.ffn(blk.norm2(x2) * (1 + ee[4]) + ee[3])
    x3 = x2 + y * ee[5]
    rel(d["x_b0"], x3, "x_b0")
    rel(got, o(**{k: v.float() for k, v in inp.items()}), "final")
