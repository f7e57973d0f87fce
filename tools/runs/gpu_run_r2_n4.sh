#!/bin/bash
# 4 GPUs: FLUX headline (SPMD), FLUX / WAN batch 1 sequence-parallel through the node API, SDXL config 5 (40/40/15/5 split)
O=gpurun_out/r2n4; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29611 bench.py --gpus 4 --steps 8 --warmup 4 > $O/spmd_n4.json 2> $O/spmd_n4.err
timeout 300 python bench.py --gpus 4 --api nodes --batch 1 --steps 8 --warmup 4 --no-bf16 > $O/nodes_b1_ulysses_n4.json 2> $O/nodes_b1_ulysses_n4.err
timeout 400 python tools/bench_wan.py --api nodes --gpus 4 --batch 1 --dtype fp8 --steps 5 --warmup 4 > $O/wan_nodes_b1_n4_fp8.json 2> $O/wan_nodes_b1_n4_fp8.err
timeout 400 $TR --master-port 29612 tools/bench_sdxl.py --gpus 4 --config 5 --steps 6 --warmup 4 > $O/sdxl5_ours_n4.json 2> $O/sdxl5_ours_n4.err
timeout 500 $TR --master-port 29613 tools/bench_sdxl.py --gpus 4 --config 5 --steps 4 --warmup 3 --impl reference > $O/sdxl5_ref_n4.json 2> $O/sdxl5_ref_n4.err
for f in spmd_n4 nodes_b1_ulysses_n4 wan_nodes_b1_n4_fp8 sdxl5_ours_n4 sdxl5_ref_n4; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json")); print("$f", d.get("ms_per_step"), d.get("e2e",{}).get("ms_per_step"), d.get("clocks",{}).get("sm_mhz"), d.get("output_matches_n1"), (d.get("bf16") or {}).get("ms_per_step"), d.get("unavailable"))
except Exception as e:
    print("$f ERR", e); print(open("$O/$f.err").read()[-1200:])
PY
done
