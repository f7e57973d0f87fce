#!/usr/bin/env python
"""Achieved time of every fused compute+collective kernel vs its roofline (BASELINE.json north star: "each fused path
reported as achieved fraction of its roofline - the slower of its compute at peak and its bytes over NVLink at link
bandwidth").  ONE process, two GPUs: the kernel runs on cuda:1, the lead's buffers live on cuda:0 (peer mapping) - what a
non-lead replica does every step - and, for comparison, with the same buffers local (NVLink exposure = difference).

Roofline terms per launch: FLOPs / measured bf16 GEMM peak, NVLink bytes / 770 GB/s (measured peer-copy bandwidth of
this pool, B200_PROFILING.md), and - because these kernels also stream their result to local HBM - local bytes /
measured HBM bandwidth.  The target is the max of the three.   python tools/fused_paths_roofline.py  (needs 2 GPUs)
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from comfyui_parallelanything_b200 import ops  # noqa: E402

NVL_GBS = 770.0


def timed(fn, dev, warm=200, iters=200):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / iters * 1e3      # us


def main() -> int:
    C_ = ops.require()
    assert torch.cuda.device_count() >= 2, "needs 2 GPUs"
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf, hbm = peaks.get("bf16_tflops", 1590.0), peaks.get("hbm_gbs", 6650.0)
    C_.enable_peer_access(1, 0)
    d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
    torch.cuda.set_device(d1)
    bf = torch.bfloat16
    rows = []

    def add(name, us_peer, us_local, flop, nvl_bytes, hbm_bytes):
        t_c, t_n, t_h = flop / (tf * 1e12) * 1e6, nvl_bytes / (NVL_GBS * 1e9) * 1e6, hbm_bytes / (hbm * 1e9) * 1e6
        target = max(t_c, t_n, t_h)
        rows.append(dict(path=name, us_peer=round(us_peer, 2), us_local=round(us_local, 2), nvlink_exposed_us=round(us_peer - us_local, 2),
                         flop=flop, nvlink_bytes=nvl_bytes, hbm_bytes=hbm_bytes, t_compute_us=round(t_c, 2),
                         t_nvlink_us=round(t_n, 2), t_hbm_us=round(t_h, 2), roofline_us=round(target, 2),
                         frac_of_roofline=round(target / us_peer, 3)))

    for n in (1, 4):
        # ---- FLUX fused scatter: peer latent shard -> patchify -> img_in GEMM (+ timestep / guidance sinusoids)
        w, b = torch.randn(3072, 64, dtype=bf, device=d1) * 0.1, torch.randn(3072, dtype=bf, device=d1)
        X = torch.empty(n, 4608, 3072, dtype=bf, device=d1)
        te, ge = (torch.empty(n, 256, dtype=bf, device=d1) for _ in range(2))
        xc = torch.empty(n, 16, 128, 128, dtype=bf, device=d1)
        us = {}
        for where, dv in (("peer", d0), ("local", d1)):
            x = torch.randn(n, 16, 128, 128, dtype=bf, device=dv)
            t = torch.rand(n, dtype=bf, device=dv)
            torch.cuda.synchronize(dv)
            us[where] = timed(lambda: C_.scatter_patch_embed(w, b, x.data_ptr(), t.data_ptr(), t.data_ptr(), te, ge, xc, X[:, 512:],
                                                              16, 128, 128, 1000.0), d1)
        add(f"flux scatter+patch-embed (n={n})", us["peer"], us["local"], 2.0 * n * 4096 * 64 * 3072, n * 16 * 128 * 128 * 2,
            n * 4096 * 3072 * 2 + n * 16 * 128 * 128 * 2)
        # ---- FLUX fused gather: final GEMM + unpatchify + Euler + peer stores into the lead's buffer
        xm = torch.randn(n, 4096, 3072, dtype=bf, device=d1)
        wf, bfin = torch.randn(64, 3072, dtype=bf, device=d1) * 0.02, torch.randn(64, dtype=bf, device=d1)
        x1 = torch.randn(n, 16, 128, 128, dtype=bf, device=d1)
        sig = torch.tensor([[1.0, 0.9]] * n, device=d1)
        us = {}
        for where, dv in (("peer", d0), ("local", d1)):
            o = torch.zeros(n, 16, 128, 128, dtype=bf, device=dv)
            torch.cuda.synchronize(dv)
            us[where] = timed(lambda: ops.gemm(xm, wf, "euler_unpatch", bias=bfin, C=16, Hl=128, Wl=128, xout_sample_off=0,
                                               x_out_ptr=o.data_ptr(), sigmas=sig, x_in=x1), d1)
        add(f"flux final GEMM + unpatchify + Euler + gather (n={n})", us["peer"], us["local"], 2.0 * n * 4096 * 3072 * 64,
            n * 16 * 128 * 128 * 2, n * 4096 * 3072 * 2 + n * 16 * 128 * 128 * 2)
    for n in (2, 4):
        # ---- SDXL fused scatter: peer NCHW latent -> im2col -> conv_in GEMM (+ timestep sinusoid)
        wp = ops.pack_conv_in_weight(torch.randn(320, 4, 3, 3, dtype=bf, device=d1) * 0.2)
        b = torch.randn(320, dtype=bf, device=d1)
        out = torch.empty(n, 128 * 128, 320, dtype=bf, device=d1)
        temb = torch.empty(n, 320, dtype=bf, device=d1)
        xc = torch.empty(n, 4, 128, 128, dtype=bf, device=d1)
        us = {}
        for where, dv in (("peer", d0), ("local", d1)):
            x = torch.randn(n, 4, 128, 128, dtype=bf, device=dv)
            t = (torch.rand(n, device=dv) * 999).to(bf)
            torch.cuda.synchronize(dv)
            us[where] = timed(lambda: C_.scatter_conv_in(wp, b, x.data_ptr(), t.data_ptr(), temb, xc, out, 4, 128, 128, 1.0, 10000.0), d1)
        add(f"sdxl scatter+conv_in (n={n})", us["peer"], us["local"], 2.0 * n * 16384 * 36 * 320, n * 4 * 128 * 128 * 2,
            n * 16384 * 320 * 2)
        # ---- SDXL fused gather: eps (NHWC) -> Euler -> NCHW peer store
        eps = torch.randn(n, 16384, 32, dtype=bf, device=d1)
        xl = torch.randn(n, 4, 128, 128, dtype=bf, device=d1)
        sig = torch.tensor([[14.6, 12.0]] * n, device=d1)
        us = {}
        for where, dv in (("peer", d0), ("local", d1)):
            o = torch.zeros(n, 4, 128, 128, dtype=bf, device=dv)
            torch.cuda.synchronize(dv)
            us[where] = timed(lambda: C_.unet_out_gather(eps, xl, o.data_ptr(), sig, n, 4, False, 1.0, 1, 0), d1)
        add(f"sdxl Euler + NCHW gather (n={n})", us["peer"], us["local"], 0.0, n * 4 * 128 * 128 * 2, n * 16384 * 32 * 2)
    out = dict(peaks=dict(bf16_tflops=tf, hbm_gbs=hbm, nvlink_gbs=NVL_GBS), rows=rows)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fused_paths_roofline.json"), "w"), indent=1)
    print(f"{'fused path':58s} {'peer us':>8s} {'local us':>8s} {'roofline us':>11s} {'(c/nvl/hbm)':>20s} {'frac':>6s}")
    for r in rows:
        print(f"{r['path']:58s} {r['us_peer']:8.2f} {r['us_local']:8.2f} {r['roofline_us']:11.2f} "
              f"{r['t_compute_us']:6.2f}/{r['t_nvlink_us']:5.2f}/{r['t_hbm_us']:6.2f} {r['frac_of_roofline']:6.3f}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
