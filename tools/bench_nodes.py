#!/usr/bin/env python
"""The ComfyUI-facing path end to end, in ONE process: ParallelDevice chain -> ParallelAnything.setup_parallel ->
hooked ``model.forward`` (what a KSampler calls) on FLUX.1-dev (random init), 1024x1024, batch 8, bf16.
Device-timed with CUDA events on the lead GPU's current stream (the hooked forward returns on that stream).

    python tools/bench_nodes.py --gpus 2 [--steps 5] [--warmup 3]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8)
    a = ap.parse_args()
    import torch
    import comfyui_parallelanything_b200 as pa
    from comfyui_parallelanything_b200.models import flux
    lead = torch.device("cuda:0")
    torch.cuda.set_device(lead)
    params = flux.flux_dev_params()
    torch.manual_seed(0)
    with torch.device(lead):
        model = flux.Flux(params, dtype=torch.bfloat16).eval()
    chain = None
    for i in range(a.gpus):
        chain = pa.ParallelDevice().add_device(f"cuda:{i}", 100.0 / a.gpus, chain)[0]
    (model,) = pa.ParallelAnything().setup_parallel(model, chain)
    eng = model._parallel_engine
    inp = flux.example_inputs(params, a.batch, 1024, 1024, 512, device=lead, dtype=torch.bfloat16)
    # per-replica device time of the shard forward (events on the stream the replica runs on)
    spans = {}
    for name, rep in eng.replicas.items():
        if not hasattr(rep, "forward_shard"):
            continue
        orig = rep.forward_shard

        def timed(*args, _orig=orig, _name=name, **kw):
            with torch.cuda.device(torch.device(_name)):
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                r = _orig(*args, **kw)
                s1.record()
                spans.setdefault(_name, []).append((s0, s1))
                return r
        rep.forward_shard = timed
    dt = -0.05

    def step():
        with torch.no_grad():
            v = model(inp["x"], inp["timesteps"], context=inp["context"], y=inp["y"], guidance=inp["guidance"])
            return inp["x"] + dt * v

    for _ in range(a.warmup):
        out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = step()
    e1.record()
    for i in range(a.gpus):
        torch.cuda.synchronize(i)
    ms = e0.elapsed_time(e1) / a.steps
    per_replica = {k: round(sum(a_.elapsed_time(b_) for a_, b_ in v[-a.steps:]) / a.steps, 2) for k, v in spans.items()}
    graphs = {k: sum(1 for g in getattr(r, "_graphs", {}).values() if not isinstance(g, str))
              for k, r in eng.replicas.items()}
    native = [bool(getattr(r, "pa_native", False)) for r in eng.replicas.values()]
    fused = any(r.get("fused") for r in eng.metrics.rows)
    print(json.dumps({"metric": "denoise-steps/sec through the ComfyUI node API (one process)", "value": round(1000.0 / ms, 4),
                      "unit": "steps/s", "ms_per_step": round(ms, 3), "n_gpus": a.gpus, "steps": a.steps,
                      "warmup": a.warmup, "native_replicas": native, "fused_scatter_gather": bool(fused),
                      "output_finite": bool(torch.isfinite(out.float()).all().item()),
                      "cuda_graphs": os.environ.get("PA_CUDA_GRAPHS", "1"), "replica_ms": per_replica,
                      "captured_graphs": graphs,
                      "last_engine_rows": [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}
                                           for r in eng.metrics.rows[-2:]],
                      "config": {"model": "FLUX.1-dev DiT 1024x1024", "global_batch": a.batch}}))
    pa.cleanup_parallel_model(model)
    return 0


if __name__ == "__main__":
    sys.exit(main())
