#!/usr/bin/env python
"""Tiny launch sets for `ncu --set full` captures of the hot kernels at FLUX shapes.
    python tools/ncu_targets.py gemm|attn|lnmod|scatter
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from comfyui_parallelanything_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
torch.manual_seed(0)
bf = dict(dtype=torch.bfloat16, device=dev)
if which == "gemm":            # 2 samples of a FLUX single-block linear2-like GEMM with gated residual
    M, K, N = 9216, 3072, 9216
    a, w, b = torch.randn(M, K, **bf), torch.randn(N, K, **bf) * 0.02, torch.randn(N, **bf)
    out = torch.empty(M, N, **bf)
    for _ in range(4):
        ops.gemm(a, w, "bias", out=out, bias=b)
elif which == "gemm_gelu":
    M, K, N = 9216, 3072, 12288
    a, w, b = torch.randn(M, K, **bf), torch.randn(N, K, **bf) * 0.02, torch.randn(N, **bf)
    out = torch.empty(M, N, **bf)
    for _ in range(4):
        ops.gemm(a, w, "gelu", out=out, bias=b)
elif which == "gemm_sdxl":    # SDXL transformer to_out / proj: small K, residual epilogue
    M, K, N = 16384, 1280, 1280
    a, w, b = torch.randn(M, K, **bf), torch.randn(N, K, **bf) * 0.03, torch.randn(N, **bf)
    res = torch.randn(M, N, **bf)
    out = torch.empty(M, N, **bf)
    for _ in range(4):
        ops.gemm(a, w, "res", out=out, bias=b, residual=res)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.gemm(a, w, "res", out=out, bias=b, residual=res)
    e1.record()
    torch.cuda.synchronize()
    print("gemm_sdxl res ms", e0.elapsed_time(e1) / 20, "TFLOP/s", 2.0 * M * N * K / (e0.elapsed_time(e1) / 20) / 1e9)
    e0.record()
    for _ in range(20):
        ops.gemm(a, w, "bias", out=out, bias=b)
    e1.record()
    torch.cuda.synchronize()
    print("gemm_sdxl bias ms", e0.elapsed_time(e1) / 20, "TFLOP/s", 2.0 * M * N * K / (e0.elapsed_time(e1) / 20) / 1e9)
elif which == "gemm_2cta":     # CTA-pair kernel at a FLUX single-block linear2-like shape
    M, K, N = 9216, 3072, 9216
    a, w, b = torch.randn(M, K, **bf), torch.randn(N, K, **bf) * 0.02, torch.randn(N, **bf)
    out = torch.empty(M, N, **bf)
    for _ in range(4):
        ops.gemm(a, w, "bias", out=out, bias=b, force_bn=512)
elif which == "gemm_2cta_gelu":   # epilogue-heavy case the eight epilogue warps were added for
    M, K, N = 9216, 3072, 12288
    a, w, b = torch.randn(M, K, **bf), torch.randn(N, K, **bf) * 0.02, torch.randn(N, **bf)
    out = torch.empty(M, N, **bf)
    for _ in range(4):
        ops.gemm(a, w, "gelu", out=out, bias=b, force_bn=512)
elif which == "attn3":
    q, k, v = (torch.randn(1, 24, 4608, 128, **bf) for _ in range(3))
    for _ in range(4):
        ops.attention(q, k, v, variant=3)
elif which == "rmsmod":
    x, w_ = torch.randn(8, 4352, 3840, **bf), torch.randn(3840, **bf)
    sc = torch.randn(8, 3840, **bf)
    o = torch.empty_like(x)
    for _ in range(4):
        ops.rmsnorm_modulate(x, o, weight=w_, scale=sc)
elif which == "attn2":
    q, k, v = (torch.randn(1, 24, 4608, 128, **bf) for _ in range(3))
    for _ in range(4):
        ops.attention(q, k, v, variant=2)
elif which == "mxfp8":
    M, K, N = 9216, 3072, 9216
    a, w = torch.randn(M, K, **bf), torch.randn(N, K, **bf) * 0.02
    aq, sfa = ops.quantize_mxfp8(a)
    wq, sfb = ops.quantize_mxfp8(w, 224)
    out = torch.empty(M, N, **bf)
    for _ in range(4):
        ops.gemm_fp8(aq, sfa, wq, sfb, "bias", 224, out=out)
elif which == "xattn":         # SDXL cross-attention (77 keys, head_dim 64, 1024 queries x 20 heads x 16 samples): CTA-pair kernel
    q = torch.randn(16, 20, 1024, 64, **bf)
    k, v = torch.randn(16, 20, 77, 64, **bf), torch.randn(16, 20, 77, 64, **bf)
    out = torch.empty(16, 1024, 20 * 64, **bf)
    for _ in range(4):
        ops.attention(q, k, v, out=out, variant=5)
elif which == "groupnorm":     # SDXL first-level GroupNorm + SiLU: 16 x 16384 x 320 (cluster of 8 CTAs per sample / channel slab)
    x = torch.randn(16, 128 * 128, 320, **bf)
    g, b_ = torch.randn(320, **bf), torch.randn(320, **bf)
    o = torch.empty_like(x)
    C_ = ops.require()
    for _ in range(4):
        C_.groupnorm_silu(x, o, g, b_, 32, 1e-5, True)
elif which == "mxfp8_256":     # 256-wide pair tiles, plain epilogue: split-N accumulators (PA_MXFP8_SPLITN=0: single accumulator)
    M, K, N = 9216, 3072, 9216
    a, w = torch.randn(M, K, **bf), torch.randn(N, K, **bf) * 0.02
    aq, sfa = ops.quantize_mxfp8(a)
    wq, sfb = ops.quantize_mxfp8(w, 256)
    out = torch.empty(M, N, **bf)
    for _ in range(4):
        ops.gemm_fp8(aq, sfa, wq, sfb, "bias", 256, out=out)
elif which == "mxfp8_l1":      # FLUX single-block linear1 (one sample): drain-first QKV+RoPE epilogue, MXFP8 MLP half
    from comfyui_parallelanything_b200.utils.selfcheck import _rope_table
    H, L, hid, mlp = 24, 4608, 3072, 12288
    a, w = torch.randn(1, L, hid, **bf), torch.randn(3 * hid + mlp, hid, **bf) * 0.02
    bias = torch.randn(3 * hid + mlp, **bf)
    aq, sfa = ops.quantize_mxfp8(a)
    wq, sfb = ops.quantize_mxfp8(w, 256)
    q, k, v = (torch.empty(1, H, L, 128, **bf) for _ in range(3))
    qs, ks = torch.ones(128, **bf), torch.ones(128, **bf)
    cat8 = torch.zeros(1, L, hid + mlp, dtype=torch.uint8, device=dev)
    cat8_sf = torch.zeros((L // 128) * ((hid + mlp) // 128) * 512, dtype=torch.uint8, device=dev)
    rope = torch.randn(L, 64, 2, device=dev)
    def l1():
        ops.gemm_fp8(aq, sfa, wq, sfb, "qkv_rope", 256, bias=bias, q=q, k=k, v=v, q_scale=qs, k_scale=ks, rope=rope,
                     seq_off=0, out8=cat8, sf8=cat8_sf, out8_col_off=hid)
    for _ in range(4):
        l1()
    if os.environ.get("PA_TIME"):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(50):
            l1()
        ev[1].record()
        torch.cuda.synchronize()
        us = ev[0].elapsed_time(ev[1]) / 50 * 1e3
        print("mxfp8_l1 us/launch", us, "TFLOP/s", 2.0 * L * (3 * hid + mlp) * hid / us / 1e6)
elif which == "scatter_conv":
    C_ = ops.require()
    x = torch.randn(2, 4, 128, 128, **bf)
    w4, b = torch.randn(320, 4, 3, 3, **bf) * 0.2, torch.randn(320, **bf)
    t = torch.rand(2, **bf) * 999
    out = torch.empty(2, 128 * 128, 320, **bf)
    temb = torch.empty(2, 320, **bf)
    xc = torch.empty_like(x)
    wp = ops.pack_conv_in_weight(w4)
    for _ in range(4):
        C_.scatter_conv_in(wp, b, x.data_ptr(), t.data_ptr(), temb, xc, out, 4, 128, 128, 1.0, 10000.0)
elif which in ("scatter_peer", "gather_peer", "scatter_conv_peer"):
    # ONE process, two GPUs: the kernel runs on cuda:1, the lead's buffers live on cuda:0 (NVLink peer mapping), exactly
    # what a non-lead replica of the in-process engine does every step.  Capture with the NVLink byte counters.
    C_ = ops.require()
    assert torch.cuda.device_count() >= 2, "needs 2 GPUs"
    C_.enable_peer_access(1, 0)
    d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
    n = 4
    torch.cuda.set_device(d1)
    if which == "scatter_peer":
        x0 = torch.randn(n, 16, 128, 128, dtype=torch.bfloat16, device=d0)
        t0 = torch.rand(n, dtype=torch.bfloat16, device=d0)
        w, b = torch.randn(3072, 64, dtype=torch.bfloat16, device=d1) * 0.1, torch.randn(3072, dtype=torch.bfloat16, device=d1)
        X = torch.empty(n, 4608, 3072, dtype=torch.bfloat16, device=d1)
        te, ge = (torch.empty(n, 256, dtype=torch.bfloat16, device=d1) for _ in range(2))
        xc = torch.empty(n, 16, 128, 128, dtype=torch.bfloat16, device=d1)
        torch.cuda.synchronize(d0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(6):
            if i == 3:
                ev[0].record()
            C_.scatter_patch_embed(w, b, x0.data_ptr(), t0.data_ptr(), t0.data_ptr(), te, ge, xc, X[:, 512:], 16, 128, 128, 1000.0)
        ev[1].record()
        torch.cuda.synchronize(d1)
        print("scatter_peer us/launch", ev[0].elapsed_time(ev[1]) / 3 * 1e3, "peer bytes/launch", x0.numel() * 2,
              "copy ok", bool(torch.equal(xc.cpu(), x0.cpu())))
    elif which == "scatter_conv_peer":
        x0 = torch.randn(n, 4, 128, 128, dtype=torch.bfloat16, device=d0)
        t0 = (torch.rand(n, device=d0) * 999).to(torch.bfloat16)
        wp = ops.pack_conv_in_weight(torch.randn(320, 4, 3, 3, dtype=torch.bfloat16, device=d1) * 0.2)
        b = torch.randn(320, dtype=torch.bfloat16, device=d1)
        out = torch.empty(n, 128 * 128, 320, dtype=torch.bfloat16, device=d1)
        temb = torch.empty(n, 320, dtype=torch.bfloat16, device=d1)
        xc = torch.empty(n, 4, 128, 128, dtype=torch.bfloat16, device=d1)
        torch.cuda.synchronize(d0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(6):
            if i == 3:
                ev[0].record()
            C_.scatter_conv_in(wp, b, x0.data_ptr(), t0.data_ptr(), temb, xc, out, 4, 128, 128, 1.0, 10000.0)
        ev[1].record()
        torch.cuda.synchronize(d1)
        print("scatter_conv_peer us/launch", ev[0].elapsed_time(ev[1]) / 3 * 1e3, "peer bytes/launch", x0.numel() * 2,
              "copy ok", bool(torch.equal(xc.cpu(), x0.cpu())))
    else:
        out0 = torch.zeros(n, 16, 128, 128, dtype=torch.bfloat16, device=d0)
        xm = torch.randn(n, 4096, 3072, dtype=torch.bfloat16, device=d1)
        wf, bfin = torch.randn(64, 3072, dtype=torch.bfloat16, device=d1) * 0.02, torch.randn(64, dtype=torch.bfloat16, device=d1)
        x1 = torch.randn(n, 16, 128, 128, dtype=torch.bfloat16, device=d1)
        sig = torch.tensor([[1.0, 0.9]] * n, device=d1)
        torch.cuda.synchronize(d0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(6):
            if i == 3:
                ev[0].record()
            ops.gemm(xm, wf, "euler_unpatch", bias=bfin, C=16, Hl=128, Wl=128, xout_sample_off=0, x_out_ptr=out0.data_ptr(),
                     sigmas=sig, x_in=x1)
        ev[1].record()
        torch.cuda.synchronize(d1)
        print("gather_peer us/launch", ev[0].elapsed_time(ev[1]) / 3 * 1e3, "peer bytes/launch", out0.numel() * 2,
              "finite", bool(torch.isfinite(out0.float()).all()))
elif which == "scatter":
    C_ = ops.require()
    x = torch.randn(2, 16, 128, 128, **bf)
    w, b = torch.randn(3072, 64, **bf) * 0.1, torch.randn(3072, **bf)
    t = torch.rand(2, **bf)
    X = torch.empty(2, 4608, 3072, **bf)
    te, ge = torch.empty(2, 256, **bf), torch.empty(2, 256, **bf)
    for _ in range(4):
        C_.scatter_patch_embed(w, b, x.data_ptr(), t.data_ptr(), t.data_ptr(), te, ge, None, X[:, 512:], 16, 128, 128, 1000.0)
elif which == "conv":
    x = torch.randn(4, 128, 128, 320, **bf)
    wt = ops.pack_conv_weight(torch.randn(320, 320, 3, 3, **bf) * 0.02)
    for _ in range(4):
        ops.conv2d_nhwc(x, wt, 9, 1, "bias")
elif which == "attn":
    q, k, v = (torch.randn(1, 24, 4608, 128, **bf) for _ in range(3))
    for _ in range(4):
        ops.attention(q, k, v)
elif which == "lnmod":
    x = torch.randn(2, 4608, 3072, **bf)
    sc, sh = torch.randn(2, 3072, **bf), torch.randn(2, 3072, **bf)
    for _ in range(4):
        ops.layernorm_modulate(x, scale=sc, shift=sh)
torch.cuda.synchronize()
print("done", which)
