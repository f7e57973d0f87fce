#!/usr/bin/env python
"""Timeline of one CTA of the ping-pong attention kernel (clock64 stamps, see PA_TR in attention2.cu).
    python tools/attn_trace.py [dbg-mask ...]      (64 = full kernel, 67 = no softmax work, 79 = also light MMAs)
Prints, per KV tile j, the cycle offsets of every event relative to the MMA thread's p_full[A] wake-up of tile j0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from comfyui_parallelanything_b200 import ops  # noqa: E402

masks = [int(a) for a in sys.argv[1:]] or [192]
dev = torch.device("cuda:0")
q, k, v = (torch.randn(2, 24, 4608, 128, dtype=torch.bfloat16, device=dev) for _ in range(3))
out = torch.empty(2, 4608, 3072, dtype=torch.bfloat16, device=dev)
C = ops.require()
names = {0: "mmaA", 1: "mmaB", 2: "smxA", 3: "smxB", 4: "tma "}
slots = {0: ["p_full", "pv_issued", "k_full", "qk_issued", "v_full"], 1: ["p_full", "pv_issued", "k_full", "qk_issued"],
         2: ["s_full", "ld_done", "max_done", "exp_done", "arrived"], 3: ["s_full", "ld_done", "max_done", "exp_done", "arrived"],
         4: ["k_empty", "v_empty"]}
for m in masks:
    for _ in range(2):
        ops.attention(q, k, v, out=out, variant=20 + m)
    torch.cuda.synchronize()
    t = C.attention2_trace()
    j0 = 10
    base = t[0, j0, 0].item()
    print(f"=== dbg mask {m}: period(j) = mmaA.p_full[j+1]-[j]: ",
          [int(t[0, j + 1, 0] - t[0, j, 0]) for j in range(4, 30)])
    for j in range(j0, j0 + 3):
        ev = []
        for r in range(5):
            for si, sn in enumerate(slots[r]):
                val = t[r, j, si].item()
                if val:
                    ev.append((val - base, f"{names[r]}.{sn}[{j}]"))
        for dt, n in sorted(ev):
            print(f"  {dt:8d}  {n}")
