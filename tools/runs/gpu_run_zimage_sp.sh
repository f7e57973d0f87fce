#!/bin/bash
# Z-Image sequence-parallel: tests (all three families), then batch 1 through the node API on 2 GPUs vs 1
O=gpurun_out/zsp; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_engine.py -x -q -k "ulysses" 2>&1 | tail -15 > $O/pytest_ulysses.log; cat $O/pytest_ulysses.log
grep -q "failed\|error" $O/pytest_ulysses.log && exit 0
for n in 2 1; do
  timeout 300 python tools/bench_zimage.py --api nodes --gpus $n --batch 1 --dtype fp8 --steps 8 --warmup 4 > $O/zimage_nodes_b1_n${n}_fp8.json 2> $O/zimage_nodes_b1_n${n}_fp8.err
  python - <<PY
import json
try:
    d=json.load(open("$O/zimage_nodes_b1_n${n}_fp8.json")); print("N=$n", d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"].get("sm_mhz"), d.get("output_matches_n1"), d["engine"].get("counters"))
except Exception as e:
    print("ERR", e); print(open("$O/zimage_nodes_b1_n${n}_fp8.err").read()[-1500:])
PY
done
timeout 300 python tools/bench_wan.py --api nodes --gpus 2 --batch 1 --dtype fp8 --steps 4 --warmup 4 > $O/wan_nodes_b1_n2_fp8.json 2> $O/wan_nodes_b1_n2_fp8.err; python -c "
import json; d=json.load(open('$O/wan_nodes_b1_n2_fp8.json')); print('wan n2', d['ms_per_step'], d['config']['parallelism'])"
