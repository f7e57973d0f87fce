#!/bin/bash
# 8 GPUs after the CTA-pair MXFP8 kernel: FLUX headline (SPMD + node API), batch-1 sequence-parallel FLUX / WAN, SDXL config 2
O=gpurun_out/r2n8b; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29711 bench.py --gpus 8 --steps 10 --warmup 5 > $O/spmd_n8.json 2> $O/spmd_n8.err
timeout 300 python bench.py --gpus 8 --api nodes --steps 10 --warmup 5 --no-bf16 > $O/nodes_b8_n8.json 2> $O/nodes_b8_n8.err
timeout 300 python bench.py --gpus 8 --api nodes --batch 1 --steps 10 --warmup 5 --no-bf16 > $O/nodes_b1_ulysses_n8.json 2> $O/nodes_b1_ulysses_n8.err
timeout 400 python tools/bench_wan.py --api nodes --gpus 8 --batch 1 --dtype fp8 --steps 6 --warmup 4 > $O/wan_nodes_b1_n8_fp8.json 2> $O/wan_nodes_b1_n8_fp8.err
timeout 300 $TR --master-port 29712 tools/bench_sdxl.py --gpus 8 --steps 10 --warmup 5 > $O/sdxl_n8.json 2> $O/sdxl_n8.err
for f in spmd_n8 nodes_b8_n8 nodes_b1_ulysses_n8 wan_nodes_b1_n8_fp8 sdxl_n8; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json")); print("$f", d.get("ms_per_step"), d.get("e2e",{}).get("ms_per_step"), d.get("clocks",{}).get("sm_mhz"), d.get("output_matches_n1"), (d.get("bf16") or {}).get("ms_per_step"))
except Exception as e:
    print("$f ERR", e); print(open("$O/$f.err").read()[-1200:])
PY
done
