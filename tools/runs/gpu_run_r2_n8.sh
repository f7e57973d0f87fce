#!/bin/bash
# round-2 measurements on 8 GPUs (one box): SPMD headline (fp8 + bf16 in one run), node API (one process) at batch 8 and
# batch 1 (sequence-parallel), fused vs NCCL scatter/gather, SDXL config 2, flag-protocol stress with 8 ranks
N=${1:-8}
O=gpurun_out/r2n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 500 $TR bench.py --gpus $N --steps 10 --warmup 5 > $O/spmd.json 2> $O/spmd.err
timeout 500 python bench.py --gpus $N --api nodes --steps 10 --warmup 5 > $O/nodes_b8.json 2> $O/nodes_b8.err
timeout 500 python bench.py --gpus $N --api nodes --batch 1 --steps 10 --warmup 5 > $O/nodes_b1_ulysses.json 2> $O/nodes_b1_ulysses.err
timeout 500 $TR bench.py --gpus $N --steps 10 --warmup 5 --backend nccl --no-bf16 > $O/spmd_nccl.json 2> $O/spmd_nccl.err
timeout 500 $TR tools/bench_sdxl.py --gpus $N --steps 10 --warmup 5 > $O/sdxl.json 2> $O/sdxl.err
timeout 300 $TR tools/flag_stress.py --epochs 30000 2>&1 | grep PA_FLAGS > $O/flag_stress.json
for f in spmd nodes_b8 nodes_b1_ulysses spmd_nccl sdxl; do echo "== $f"; python - <<PY
import json
try:
    d = json.load(open("$O/$f.json"))
    print(d.get("ms_per_step"), d.get("e2e", {}).get("ms_per_step"), d.get("dtype", "")[:5], "bf16:", d.get("bf16", {}).get("ms_per_step"), d.get("output_matches_n1"), d.get("clocks", {}).get("sm_mhz"), (d.get("engine") or {}).get("setup"), (d.get("engine") or {}).get("counters"))
except Exception as e:
    print("ERR", e); print(open("$O/$f.err").read()[-1200:])
PY
done
cat $O/flag_stress.json
