#!/usr/bin/env python
"""Per-kernel device-time breakdown of one FLUX denoise step on the native executor.

Every native op is bracketed by CUDA events (and a synchronize, so launches are serialised — the
SHARES are what matters, the sum is a slight over-estimate of the real step).  GEMMs are grouped by
(mode, M, N, K) with achieved TFLOP/s against the measured cuBLAS peak in MEASURED_PEAKS.json.

    python tools/profile_flux.py --batch 8 [--steps 2] -> gpurun_out/profile_flux.json + table
"""
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
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--fp8", action="store_true", help="MXFP8 block GEMMs (the bench default)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "profile_flux.json"))
    a = ap.parse_args()
    import torch
    from comfyui_parallelanything_b200 import ops
    from comfyui_parallelanything_b200.exec import flux_exec
    from comfyui_parallelanything_b200.models import flux
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    params = flux.flux_dev_params()
    torch.manual_seed(0)
    with torch.device(dev):
        model = flux.Flux(params, dtype=torch.bfloat16)
    ex = flux_exec.FluxExecutor(model, dev, fp8=a.fp8)
    ex.dual_stream = False          # serialised per-op timing: keep one stream
    del model
    inp = flux.example_inputs(params, a.batch, 1024, 1024, 512, device=dev, dtype=torch.bfloat16)
    sig = torch.tensor([[1.0, 0.9]] * a.batch, device=dev)
    x, t, c, y, g = ex._prep(inp["x"], inp["timesteps"], inp["context"], inp["y"], inp["guidance"])
    ex.denoise_step(x, t, c, y, g, sig)           # warm-up (workspaces, attributes)
    torch.cuda.synchronize()

    rec = collections.OrderedDict()
    recording = [False]

    def wrap(name, fn, keyf, flopf):
        def inner(*args, **kw):
            if not recording[0]:
                return fn(*args, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*args, **kw)
            e1.record()
            e1.synchronize()
            k = (name,) + tuple(keyf(*args, **kw))
            ent = rec.setdefault(k, dict(ms=0.0, n=0, flop=0.0, bytes=0.0))
            ent["ms"] += e0.elapsed_time(e1)
            ent["n"] += 1
            fl, by = flopf(*args, **kw)
            ent["flop"] += fl
            ent["bytes"] += by
            return r
        return inner

    def gemm_key(a_, w, mode="bias", **kw):
        M = a_.numel() // a_.shape[-1]
        return (mode, M, w.shape[0], w.shape[1])

    def gemm_flop(a_, w, mode="bias", **kw):
        M = a_.numel() // a_.shape[-1]
        N, K = w.shape
        return 2.0 * M * N * K, 2.0 * (M * K + N * K + M * N)

    def attn_key(q, k, v, out=None, scale=None):
        return tuple(q.shape)

    def attn_flop(q, k, v, out=None, scale=None):
        b, h, l, d = q.shape
        return 4.0 * b * h * l * k.shape[2] * d, 2.0 * 4 * q.numel()

    def ln_key(x_, out=None, **kw):
        return tuple(x_.shape)

    def ln_flop(x_, out=None, **kw):
        return 0.0, 4.0 * x_.numel()

    def fp8_key(aq, sfa, wq, sfb, mode="bias", w_tile=224, **kw):
        M = aq.numel() // aq.shape[-1]
        return (mode, M, wq.shape[0], wq.shape[1])

    def fp8_flop(aq, sfa, wq, sfb, mode="bias", w_tile=224, **kw):
        M = aq.numel() // aq.shape[-1]
        N, K = wq.shape
        return 2.0 * M * N * K, 1.0 * (M * K + N * K) + 2.0 * M * N

    def q_key(x_, tile_rows=128):
        return tuple(x_.shape)

    def q_flop(x_, tile_rows=128):
        return 0.0, 3.0 * x_.numel()

    C = ops.require()
    ops.gemm = wrap("gemm", ops.gemm, gemm_key, gemm_flop)
    ops.gemm_fp8 = wrap("gemm_fp8", ops.gemm_fp8, fp8_key, fp8_flop)
    ops.quantize_mxfp8 = wrap("quantize_mxfp8", ops.quantize_mxfp8, q_key, q_flop)
    ops.attention = wrap("attention", ops.attention, attn_key, attn_flop)
    ops.layernorm_modulate = wrap("ln_mod", ops.layernorm_modulate, ln_key, ln_flop)
    orig_scatter = C.scatter_patch_embed

    recording[0] = True
    tot0, tot1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(a.steps):
        ex.denoise_step(x, t, c, y, g, sig)
    torch.cuda.synchronize()
    recording[0] = False
    # un-instrumented step time for reference
    tot0.record()
    for _ in range(a.steps):
        ex.denoise_step(x, t, c, y, g, sig)
    tot1.record()
    torch.cuda.synchronize()
    step_ms = tot0.elapsed_time(tot1) / a.steps

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_bw = peaks.get("hbm_gbs", 6650.0)
    rows = []
    total = sum(v["ms"] for v in rec.values()) / a.steps
    for k, v in rec.items():
        ms = v["ms"] / a.steps
        tf = v["flop"] / a.steps / (ms * 1e-3) / 1e12 if ms > 0 else 0
        gbs = v["bytes"] / a.steps / (ms * 1e-3) / 1e9 if ms > 0 else 0
        rows.append(dict(op=" ".join(str(s) for s in k), launches=v["n"] // a.steps, ms=round(ms, 3),
                         share=round(ms / total, 4), tflops=round(tf, 1), frac_of_measured_bf16=round(tf / peak_tf, 3),
                         gbs=round(gbs, 1), frac_of_measured_hbm=round(gbs / peak_bw, 3)))
    rows.sort(key=lambda r: -r["ms"])
    out = dict(batch=a.batch, step_ms_uninstrumented=round(step_ms, 3), sum_kernel_ms=round(total, 3),
               launches_per_step=ex.launches_per_step, peak_tflops_measured_sustained=peak_tf, rows=rows)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print(f"step {step_ms:.2f} ms un-instrumented; sum of kernels {total:.2f} ms; {ex.launches_per_step} launches")
    print(f"{'op':58s} {'n':>4s} {'ms':>9s} {'share':>6s} {'TFLOP/s':>8s} {'%peak':>6s} {'GB/s':>8s}")
    for r in rows:
        print(f"{r['op']:58s} {r['launches']:4d} {r['ms']:9.3f} {r['share']*100:5.1f}% {r['tflops']:8.1f} "
              f"{r['frac_of_measured_bf16']*100:5.1f}% {r['gbs']:8.1f}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
