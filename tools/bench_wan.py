#!/usr/bin/env python
"""Secondary benchmark: WAN2.2-A14B-sized video DiT (14.3 B params, random init), 720p x 16 frames (latent
[16, 4, 90, 160] -> 14 400 tokens per sample), batch 8, bf16 - BASELINE.json config 4.  Same protocol and JSON line as
bench.py (one step = forward + Euler update; device-timed, max over ranks; e2e with per-step H2D/D2H).

    python tools/bench_wan.py --gpus 1 [--impl reference] [--batch 8]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_wan.py --gpus 8
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as hb  # noqa: E402  (helpers: ClockSampler, timed, emit, dist_env, ...)

MODEL_NAME = "WAN2.2-A14B video DiT 720p x 16 frames"


def host_inputs(B, cap_len):
    import torch
    from comfyui_parallelanything_b200.models import wan
    cfg = wan.wan22_a14b_params()
    inp = wan.example_inputs(cfg, B, frames=16, height=720, width=1280, device="cpu", dtype=torch.bfloat16)
    inp["sig"] = torch.tensor([[1.0, 0.875]] * B, dtype=torch.float32)
    return cfg, {k: v.pin_memory() for k, v in inp.items()}


def run_nodes(a) -> int:
    """One process through the node API (tools/_nodes_bench.py); batch 1 = the video workload: sequence-parallel."""
    import torch
    import _nodes_bench
    from comfyui_parallelanything_b200.models import wan
    cfg, host = host_inputs(a.batch, a.cap_len)

    def build(lead):
        with torch.device(lead):
            return wan.WanModel(cfg, dtype=torch.bfloat16).eval()

    return _nodes_bench.run(a, hb, MODEL_NAME, build, host,
                            {"model": MODEL_NAME, "global_batch": a.batch, "text_len": a.cap_len, "baseline_config": 4})


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp8"], help="fp8 = MXFP8 block GEMMs in the native executor")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--cap-len", type=int, default=512)
    ap.add_argument("--api", default="spmd", choices=["spmd", "nodes"],
                    help="nodes = ONE process through the ComfyUI node API, new latent tensor every step (batch 1 -> "
                         "sequence-parallel over all GPUs)")
    a = ap.parse_args()
    if a.api == "nodes":
        return run_nodes(a)
    hb.quiet_stdout()
    import torch
    rank, world, local = hb.dist_env()
    B = a.batch
    pcts = [100.0 / a.gpus] * a.gpus
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
    from comfyui_parallelanything_b200.models import wan
    cfg, host = host_inputs(B, a.cap_len)
    result_host = torch.empty(B, 16, 4, 90, 160, dtype=torch.bfloat16).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = result_host.numel() * result_host.element_size()
    launches = 0
    eng = None
    if a.impl == "ours":
        from comfyui_parallelanything_b200.exec.wan_exec import WanExecutor
        torch.manual_seed(1234)
        with torch.device(dev):
            model = wan.WanModel(cfg, dtype=torch.bfloat16).eval()
        ex = WanExecutor(model, dev, cuda_graphs=True, fp8=(a.dtype == "fp8"))
        del model
        torch.cuda.empty_cache()
        if world == 1:
            d = {k: v.to(dev) for k, v in host.items()}
            stage = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
            out_buf = torch.empty(B, 16, 4, 90, 160, dtype=torch.bfloat16, device=dev)

            def step_device():
                ex.denoise_step(d["x"], d["timesteps"], d["context"], d["sig"], out=out_buf)

            def step_e2e():
                for k in stage:
                    stage[k].copy_(host[k], non_blocking=True)
                ex.invalidate_conditioning()        # the staged conditioning is re-sent every step: recompute it
                ex.denoise_step(stage["x"], stage["timesteps"], stage["context"], stage["sig"], out=out_buf)
                result_host.copy_(out_buf, non_blocking=True)
        else:
            from comfyui_parallelanything_b200.parallel.spmd import SpmdWanEngine
            eng = SpmdWanEngine(ex, B, 4, 720, 1280, a.cap_len, weights=pcts)
            order = [host[k] for k in ("x", "timesteps", "context", "sig")]
            if rank == 0:
                eng.stage_inputs(*order)
            torch.cuda.synchronize()

            def step_device():
                eng.step()

            def step_e2e():
                if rank == 0:
                    eng.stage_inputs(*order)
                eng.new_conditioning()
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
            with torch.device(lead):
                model = wan.WanModel(cfg, dtype=torch.bfloat16).eval()
            torch.nn.Linear.reset_parameters = lambda self: None      # clones are overwritten anyway (see bench.py)
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
                return x + (sig[:, 1] - sig[:, 0]).view(-1, 1, 1, 1, 1).to(x.dtype) * e

            def step_device():
                with torch.no_grad():
                    e = model(d["x"], d["timesteps"], context=d["context"])
                    return euler(d["x"], e, d["sig"])

            def step_e2e():
                for k in stage:
                    stage[k].copy_(host[k], non_blocking=True)
                with torch.no_grad():
                    e = model(stage["x"], stage["timesteps"], context=stage["context"])
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
                 "vs_baseline": None, "sec_per_it": round(ms / 1000.0, 4), "dtype": (a.dtype if a.impl == "ours" else "bf16"), "data": "synthetic, random-init weights", "impl": a.impl,
                 "clocks": clocks,
                 "e2e": {"value": round(1000.0 / ms_e2e, 4), "unit": "steps/s", "ms_per_step": round(ms_e2e, 3),
                         "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                 "output_finite": bool(torch.isfinite(result_host.float()).all().item()),
                 "config": {"model": MODEL_NAME, "global_batch": B, "text_len": a.cap_len, "baseline_config": 4, "parallelism": par}})
    return 0


if __name__ == "__main__":
    sys.exit(main())
