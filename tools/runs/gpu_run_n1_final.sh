#!/bin/bash
O=gpurun_out/n1fin; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/pytest_gpu_1gpu.log; cat $O/pytest_gpu_1gpu.log
for dt in fp8 bf16; do
timeout 400 python tools/bench_wan.py --api nodes --gpus 1 --batch 1 --dtype $dt --steps 5 --warmup 4 > $O/wan_nodes_b1_n1_$dt.json 2> $O/wan_nodes_b1_n1_$dt.err
python - <<PY
import json
try:
    d=json.load(open("$O/wan_nodes_b1_n1_$dt.json")); print("$dt", d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"].get("sm_mhz"))
except Exception as e:
    print("ERR", e); print(open("$O/wan_nodes_b1_n1_$dt.err").read()[-1500:])
PY
done
