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
