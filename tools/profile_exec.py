#!/usr/bin/env python
"""Per-op device-time breakdown of one native executor step (SDXL UNet / WAN), CUDA events around every op
(serialised; shares matter).   python tools/profile_exec.py --model sdxl --batch 16"""
from __future__ import annotations

import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="sdxl", choices=["sdxl", "wan"])
    ap.add_argument("--batch", type=int, default=16)
    a = ap.parse_args()
    import torch
    from comfyui_parallelanything_b200 import ops
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    if a.model == "sdxl":
        from comfyui_parallelanything_b200.exec.unet_exec import UNetExecutor
        from comfyui_parallelanything_b200.models import unet
        cfg = unet.sdxl_config()
        m = unet.UNetModel(**cfg).to(device=dev, dtype=torch.bfloat16).eval()
        ex = UNetExecutor(m, dev)
        del m
        inp = unet.example_inputs(cfg, a.batch, 1024, 1024, 77, device=dev, dtype=torch.bfloat16)
        sig = torch.tensor([[14.6, 12.0]] * a.batch, device=dev)
        step = lambda: ex.denoise_step(inp["x"], inp["timesteps"], inp["context"], inp["y"], sig)  # noqa: E731
    else:
        from comfyui_parallelanything_b200.exec.wan_exec import WanExecutor
        from comfyui_parallelanything_b200.models import wan
        p = wan.wan22_a14b_params()
        with torch.device(dev):
            m = wan.WanModel(p, dtype=torch.bfloat16).eval()
        ex = WanExecutor(m, dev)
        del m
        inp = wan.example_inputs(p, a.batch, 16, 720, 1280, device=dev, dtype=torch.bfloat16)
        sig = torch.tensor([[1.0, 0.9]] * a.batch, device=dev)
        x, t, c = ex._prep(inp["x"], inp["timesteps"], inp["context"])
        step = lambda: ex.denoise_step(x, t, c, sig)  # noqa: E731
    step()
    torch.cuda.synchronize()
    rec = collections.OrderedDict()
    on = [False]

    def wrap(name, fn, keyf):
        def inner(*args, **kw):
            if not on[0]:
                return fn(*args, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*args, **kw)
            e1.record()
            e1.synchronize()
            k = (name,) + tuple(keyf(*args, **kw))
            ent = rec.setdefault(k, [0.0, 0])
            ent[0] += e0.elapsed_time(e1)
            ent[1] += 1
            return r
        return inner

    shp = lambda t: tuple(t.shape)  # noqa: E731
    ops.gemm = wrap("gemm", ops.gemm, lambda a_, w, mode="bias", **kw: (mode, a_.numel() // a_.shape[-1], w.shape[0], w.shape[1]))
    ops.conv2d_nhwc = wrap("conv", ops.conv2d_nhwc, lambda x, w, taps, stride=1, mode="bias", **kw: (mode, shp(x), w.shape[0], taps, stride))
    ops.attention = wrap("attention", ops.attention, lambda q, k, v, **kw: (shp(q), k.shape[2]))
    ops.groupnorm_silu = wrap("groupnorm", ops.groupnorm_silu, lambda x, *r, **kw: (shp(x),))
    ops.layernorm_modulate = wrap("layernorm", ops.layernorm_modulate, lambda x, *r, **kw: (shp(x),))
    C = ops.require()

    class CW:
        def __getattr__(self, n):
            f = getattr(C, n)
            return wrap("C." + n, f, lambda *args, **kw: ()) if callable(f) else f
    ops.require = lambda: CW()
    import comfyui_parallelanything_b200.exec.unet_exec as ue
    ue.ops = ops
    on[0] = True
    step()
    torch.cuda.synchronize()
    on[0] = False
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    step()
    e1.record()
    torch.cuda.synchronize()
    total = sum(v[0] for v in rec.values())
    rows = sorted(([" ".join(map(str, k)), v[1], v[0]] for k, v in rec.items()), key=lambda r: -r[2])
    print(f"{a.model} batch {a.batch}: step {e0.elapsed_time(e1) / 2:.2f} ms un-instrumented, kernels sum {total:.2f} ms, "
          f"{sum(v[1] for v in rec.values())} launches")
    agg = collections.defaultdict(float)
    for k, v in rec.items():
        agg[k[0]] += v[0]
    print("by op:", {k: round(v, 2) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])})
    for r in rows[:28]:
        print(f"{r[0][:92]:92s} {r[1]:4d} {r[2]:8.3f} ms {100 * r[2] / total:5.1f}%")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(model=a.model, batch=a.batch, rows=rows), open(os.path.join(ROOT, "gpurun_out", f"profile_{a.model}.json"), "w"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
