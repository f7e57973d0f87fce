#!/usr/bin/env python
"""BASELINE.json config 1 (plumbing, no GPU): SD1.5 UNet (0.86 B params, random init), 512x512, batch 2, fp32, two
CPU "devices" (``cpu``, ``cpu`` - the reference keys replicas by device string, so both chain entries share one replica
and the batch is split 1 + 1 across two worker threads).  Wall-clock seconds per denoise step (forward + Euler update),
ours vs the unmodified reference, both through their node API.

    python tools/bench_cpu_config1.py [--impl ours|reference] [--steps 2] [--warmup 1]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    import torch
    from comfyui_parallelanything_b200.models import unet
    torch.manual_seed(0)
    cfg = unet.sd15_config()
    model = unet.UNetModel(**cfg).eval()
    inp = unet.example_inputs(cfg, 2, 512, 512, ctx_len=77, device="cpu", dtype=torch.float32)
    if a.impl == "ours":
        import comfyui_parallelanything_b200 as pa
        chain = pa.ParallelDevice().add_device("cpu", 50.0, pa.ParallelDevice().add_device("cpu", 50.0, None)[0])[0]
        (model,) = pa.ParallelAnything().setup_parallel(model, chain)
    else:
        from baseline import ref_loader
        ref = ref_loader.load()
        chain = ref.ParallelDevice().add_device("cpu", 50.0, ref.ParallelDevice().add_device("cpu", 50.0, None)[0])[0]
        (model,) = ref.ParallelAnything().setup_parallel(model, chain, True, False, True, False)

    def step():
        with torch.no_grad():
            eps = model(inp["x"], inp["timesteps"], context=inp["context"])
            return inp["x"] - 0.05 * eps

    for _ in range(a.warmup):
        out = step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        out = step()
    sec = (time.perf_counter() - t0) / a.steps
    print(json.dumps({"metric": "sec/it (wall clock, CPU)", "impl": a.impl, "value": round(sec, 3), "steps": a.steps,
                      "warmup": a.warmup, "threads": torch.get_num_threads(), "output_finite": bool(torch.isfinite(out).all()),
                      "config": {"model": "SD1.5 UNet 512x512", "global_batch": 2, "devices": ["cpu", "cpu"],
                                 "baseline_config": 1}}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
