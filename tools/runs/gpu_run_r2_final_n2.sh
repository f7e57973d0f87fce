#!/bin/bash
# final 2-GPU validation of the tree: full GPU test suite + the 2-GPU headline (one process per GPU)
O=gpurun_out/r2fin; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_gpu_2gpus.log; cat $O/pytest_gpu_2gpus.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --steps 8 --warmup 4 > $O/spmd_n2.json 2> $O/spmd_n2.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2fin/spmd_n2.json")); print("spmd n2", d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"], d.get("output_matches_n1"), d["bf16"]["ms_per_step"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2fin/spmd_n2.err").read()[-1500:])
PY
