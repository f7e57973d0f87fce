#!/bin/bash
mkdir -p gpurun_out
for fam in unet wan flux; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/spmd_check.py $fam > gpurun_out/spmd2_$fam.log 2>&1
  grep "PA_SPMD" gpurun_out/spmd2_$fam.log | cut -c1-600 || tail -5 gpurun_out/spmd2_$fam.log
done
