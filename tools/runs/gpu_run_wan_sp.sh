#!/bin/bash
# WAN2.2-A14B, 720p x 16 frames, batch 1 through the node API: N GPUs sequence-parallel
N=${1:-2}; O=gpurun_out/wansp; mkdir -p $O
for dt in fp8 bf16; do
  timeout 600 python tools/bench_wan.py --api nodes --gpus $N --batch 1 --dtype $dt --steps 5 --warmup 4 > $O/wan_nodes_b1_n${N}_$dt.json 2> $O/wan_nodes_b1_n${N}_$dt.err
  python - <<PY
import json
try:
    d=json.load(open("$O/wan_nodes_b1_n${N}_$dt.json")); print("$dt N=$N", d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"].get("sm_mhz"), d.get("output_matches_n1"), d["engine"].get("counters"))
except Exception as e:
    print("ERR", e); print(open("$O/wan_nodes_b1_n${N}_$dt.err").read()[-1500:])
PY
done
