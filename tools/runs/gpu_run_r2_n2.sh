#!/bin/bash
# round-2 evidence on 2 GPUs: NVLink-counter ncu captures of the fused scatter / gather kernels (one process, kernel on
# cuda:1, lead buffers on cuda:0), fused vs NCCL A/B, weight replication (NVLS multicast vs NCCL) at full FLUX size
O=gpurun_out/r2n2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -k "groupnorm or unet" 2>&1 | tail -3 > $O/pytest_unet.log; cat $O/pytest_unet.log
for t in scatter_peer gather_peer scatter_conv_peer; do
  case $t in scatter_peer) K=scatter_patch_embed_kernel;; gather_peer) K=gemm_bf16_tcgen05_kernel;; *) K=scatter_conv_in_kernel;; esac
  timeout 300 ncu --set full --section Nvlink --section Nvlink_Tables --clock-control none --import-source on -c 1 --launch-skip 2 \
     -k regex:$K -o $O/ncu_$t -f python tools/ncu_targets.py $t > $O/ncu_$t.log 2>&1
  tail -1 $O/ncu_$t.log
  timeout 100 python tools/ncu_targets.py $t 2>&1 | tail -1 > $O/time_$t.txt; cat $O/time_$t.txt
done
timeout 400 $TR bench.py --gpus 2 --steps 6 --warmup 4 > $O/spmd_fused_n2.json 2> $O/spmd_fused_n2.err
timeout 400 $TR bench.py --gpus 2 --steps 6 --warmup 4 --backend nccl --no-bf16 > $O/spmd_nccl_n2.json 2> $O/spmd_nccl_n2.err
timeout 400 $TR bench.py --gpus 2 --steps 3 --warmup 3 --replicate nvls --no-bf16 --dtype bf16 > $O/replicate_nvls_n2.json 2> $O/replicate_nvls_n2.err
timeout 400 $TR bench.py --gpus 2 --steps 3 --warmup 3 --replicate nccl --no-bf16 --dtype bf16 > $O/replicate_nccl_n2.json 2> $O/replicate_nccl_n2.err
timeout 400 python bench.py --gpus 2 --api nodes --steps 6 --warmup 4 > $O/nodes_n2.json 2> $O/nodes_n2.err
timeout 400 $TR tools/bench_sdxl.py --gpus 2 --steps 6 --warmup 4 > $O/sdxl_n2.json 2> $O/sdxl_n2.err
for f in spmd_fused_n2 spmd_nccl_n2 replicate_nvls_n2 replicate_nccl_n2 nodes_n2 sdxl_n2; do echo "== $f"; python - <<PY
import json
try:
    d = json.load(open("$O/$f.json"))
    print(d.get("ms_per_step"), d.get("e2e", {}).get("ms_per_step"), d.get("dtype", "")[:5], d.get("setup"), d.get("output_matches_n1"), d.get("bf16", {}).get("ms_per_step"), (d.get("engine") or {}).get("setup"))
except Exception as e:
    print("ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
