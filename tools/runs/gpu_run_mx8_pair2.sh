#!/bin/bash
# CTA-pair MXFP8: per-kernel breakdown of the fp8 step + ncu of linear1 (QKV epilogue) and a 224-wide GEMM
O=gpurun_out/mx8pair; mkdir -p $O
timeout 300 python tools/profile_flux.py --fp8 --out $O/profile_flux_fp8_b8.json > $O/profile_flux_fp8_b8.txt 2>&1; tail -32 $O/profile_flux_fp8_b8.txt
for t in mxfp8_l1 mxfp8; do
  timeout 300 ncu --set full --clock-control none --import-source on -c 1 --launch-skip 2 -k regex:gemm_mxfp8 \
     -o $O/ncu_${t}_pair -f python tools/ncu_targets.py $t > $O/ncu_$t.log 2>&1
  tail -2 $O/ncu_$t.log
done
