#!/usr/bin/env python
"""Secondary benchmark: SDXL-base UNet (2.57 B params, random init) 1024x1024, bf16 — BASELINE.json configs
2 (batch 16, even split) and 5 (batch 32 on 4 GPUs, 40/40/15/5 -> reference-compatible 12/12/4/4 split).
Same protocol and JSON line as bench.py (one step = eps forward + Euler update; device-timed, max over ranks;
e2e with per-step H2D/D2H).

    python tools/bench_sdxl.py --gpus 1 --batch 16
    python -m torch.distributed.run --nproc-per-node 4 ... tools/bench_sdxl.py --gpus 4 --config 5 [--impl reference]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as hb  # noqa: E402  (helpers: ClockSampler, timed, emit, dist_env, ...)

MODEL_NAME = "SDXL-base UNet 1024x1024"


def host_inputs(B):
    import torch
    from comfyui_parallelanything_b200.models import unet
    cfg = unet.sdxl_config()
    inp = unet.example_inputs(cfg, B, 1024, 1024, ctx_len=77, device="cpu", dtype=torch.bfloat16)
    inp["sig"] = torch.tensor([[14.6, 12.0]] * B, dtype=torch.float32)
    return cfg, {k: v.pin_memory() for k, v in inp.items()}


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 5])
    ap.add_argument("--batch", type=int, default=0)
    a = ap.parse_args()
    hb.quiet_stdout()
    import torch
    rank, world, local = hb.dist_env()
    B = a.batch or (16 if a.config == 2 else 32)
    pcts = [40, 40, 15, 5][:a.gpus] if a.config == 5 and a.gpus == 4 else [100.0 / a.gpus] * a.gpus
    tworld = world
    if a.impl == "reference":
        # the reference is one process driving all GPUs: the other torchrun ranks stay off the GPUs (host-only gloo
        # barrier, no CUDA context / NCCL kernels on the devices its replicas run on) - see bench.py
        if hb.host_only_group(rank, world):
            return 0
        tworld = 1
    torch.cuda.set_device(local if a.impl == "ours" else 0)
    dev = torch.device("cuda", local)
    if world > 1 and a.impl == "ours":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from comfyui_parallelanything_b200.models import unet
    cfg, host = host_inputs(B)
    result_host = torch.empty(B, 4, 128, 128, dtype=torch.bfloat16).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = result_host.numel() * result_host.element_size()
    launches = 0
    eng = None
    if a.impl == "ours":
        from comfyui_parallelanything_b200.exec.unet_exec import UNetExecutor
        torch.manual_seed(1234)
        model = unet.UNetModel(**cfg).to(device=dev, dtype=torch.bfloat16).eval()
        ex = UNetExecutor(model, dev, cuda_graphs=True)
        del model
        if world == 1:
            d = {k: v.to(dev) for k, v in host.items()}
            stage = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
            out_buf = torch.empty(B, 4, 128, 128, dtype=torch.bfloat16, device=dev)

            def step_device():
                ex.denoise_step(d["x"], d["timesteps"], d["context"], d["y"], d["sig"], out=out_buf)

            def step_e2e():
                for k in stage:
                    stage[k].copy_(host[k], non_blocking=True)
                ex.denoise_step(stage["x"], stage["timesteps"], stage["context"], stage["y"], stage["sig"], out=out_buf)
                result_host.copy_(out_buf, non_blocking=True)
        else:
            from comfyui_parallelanything_b200.parallel.spmd import SpmdUNetEngine
            eng = SpmdUNetEngine(ex, B, 1024, 1024, 77, weights=pcts)
            order = [host[k] for k in ("x", "timesteps", "context", "y", "sig")]
            if rank == 0:
                eng.stage_inputs(*order)
            torch.cuda.synchronize()

            def step_device():
                eng.step()

            def step_e2e():
                if rank == 0:
                    eng.stage_inputs(*order)
                out = eng.step()
                if rank == 0:
                    result_host.copy_(out, non_blocking=True)
        par = f"dp{world} split {eng.sizes if eng else [B]} (fused in-kernel NVLink scatter/gather)"
    else:
        from baseline import ref_loader
        try:
            ref = ref_loader.load()
        except Exception as e:
            if rank == 0:
                hb.emit({"impl": "reference", "unavailable": str(e)[:300]})
            return 0
        if rank == 0:
            lead = torch.device("cuda", 0)
            torch.set_default_dtype(torch.bfloat16)
            torch.manual_seed(1234)
            model = unet.UNetModel(**cfg).to(device=lead, dtype=torch.bfloat16).eval()
            torch.nn.Linear.reset_parameters = lambda self: None      # clones are overwritten anyway (see bench.py)
            torch.nn.Conv2d.reset_parameters = lambda self: None
            chain = None
            for i in range(a.gpus):
                chain = ref.ParallelDevice().add_device(f"cuda:{i}", float(pcts[i]), chain)[0]
            (model,) = ref.ParallelAnything().setup_parallel(model, chain, True, False, True, False)
            with torch.no_grad():
                for t_ in list(model.parameters()) + list(model.buffers()):
                    if t_.device.type == "cpu":
                        t_.data = t_.data.to(lead)
            d = {k: v.to(lead) for k, v in host.items()}
            stage = {k: torch.empty_like(v, device=lead) for k, v in host.items()}

            def euler(x, e, sig):
                return x + (sig[:, 1] - sig[:, 0]).view(-1, 1, 1, 1).to(x.dtype) * e

            def step_device():
                with torch.no_grad():
                    e = model(d["x"], d["timesteps"], context=d["context"], y=d["y"])
                    return euler(d["x"], e, d["sig"])

            def step_e2e():
                for k in stage:
                    stage[k].copy_(host[k], non_blocking=True)
                with torch.no_grad():
                    e = model(stage["x"], stage["timesteps"], context=stage["context"], y=stage["y"])
                    result_host.copy_(euler(stage["x"], e, stage["sig"]), non_blocking=True)
        else:
            def step_device():
                return None
            step_e2e = step_device
        par = f"reference threads x{a.gpus} pct {pcts}"
    sampler = hb.ClockSampler()
    if rank == 0:
        sampler.start()
    ms = hb.timed(step_device, a.steps, a.warmup, tworld)
    ms_e2e = hb.timed(step_e2e, a.steps, max(3, a.warmup // 2), tworld)
    clocks = sampler.stop(a.gpus) if rank == 0 else {}
    if eng is not None:
        eng.check_error()
        eng.close()
    if a.impl == "reference":
        hb.host_only_release(world)
    elif world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if rank == 0:
        hb.emit({"metric": hb.METRIC, "value": round(1000.0 / ms, 4), "unit": "steps/s", "n_gpus": a.gpus, "steps": a.steps,
                 "warmup": a.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong",
                 "vs_baseline": None, "dtype": "bf16", "data": "synthetic, random-init weights", "impl": a.impl,
                 "clocks": clocks,
                 "e2e": {"value": round(1000.0 / ms_e2e, 4), "unit": "steps/s", "ms_per_step": round(ms_e2e, 3),
                         "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                 "output_finite": bool(torch.isfinite(result_host.float()).all().item()),
                 "config": {"model": MODEL_NAME, "global_batch": B, "baseline_config": a.config, "parallelism": par}})
    return 0


if __name__ == "__main__":
    sys.exit(main())
