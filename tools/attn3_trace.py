#!/usr/bin/env python
"""Timeline of one CTA pair of the CTA-pair attention kernel (PA_TR3 stamps in attention3.cu).
    PA_ATTN3_TRACE=1 python tools/attn3_trace.py"""
import os
import sys

os.environ["PA_ATTN3_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from comfyui_parallelanything_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
q, k, v = (torch.randn(2, 24, 4608, 128, dtype=torch.bfloat16, device=dev) for _ in range(3))
out = torch.empty(2, 4608, 3072, dtype=torch.bfloat16, device=dev)
C = ops.require()
names = {0: "mmaA", 1: "mmaB", 2: "smxA", 3: "smxB", 4: "tma ", 5: "peerA", 6: "peerB"}
slots = {0: ["k_full", "s_free", "qk_issued", "p_full", "p_hi", "pv_issued"],
         1: ["v_full", "s_free", "qk_issued", "p_full", "p_hi", "pv_issued"],
         2: ["s_full", "ld_done", "max_done", "p_empty", "half_stored", "all_stored"], 4: ["k_empty", "v_empty"]}
slots[3] = slots[5] = slots[6] = slots[2]
for _ in range(2):
    ops.attention(q, k, v, out=out, variant=3)
torch.cuda.synchronize()
t = C.attention3_trace()
j0 = 10
base = t[0, j0, 0].item()
print("period(j) = mmaA.k_full[j+1]-[j]:", [int(t[0, j + 1, 0] - t[0, j, 0]) for j in range(4, 30)])
for j in range(j0, j0 + 2):
    ev = []
    for r in range(5):                      # leader-CTA clocks only
        for si, sn in enumerate(slots[r]):
            val = t[r, j, si].item()
            if val:
                ev.append((val - base, f"{names[r]}.{sn}[{j}]"))
    for dt, n in sorted(ev):
        print(f"  {dt:8d}  {n}")
pb = t[5, j0, 0].item()
print("peer CTA (own clock, relative to its smxA.s_full):")
ev = []
for r in (5, 6):
    for si, sn in enumerate(slots[r]):
        for j in (j0, j0 + 1):
            val = t[r, j, si].item()
            if val:
                ev.append((val - pb, f"{names[r]}.{sn}[{j}]"))
for dt, n in sorted(ev):
    print(f"  {dt:8d}  {n}")
