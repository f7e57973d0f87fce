#!/bin/bash
# compute-sanitizer memcheck + synccheck over the round-2 kernels' self-checks (small shapes)
O=gpurun_out/sanit; mkdir -p $O
for spec in "memcheck =gemm_mxfp8" "synccheck =gemm_mxfp8" "memcheck groupnorm_cluster" "memcheck cross_attention_cluster" "memcheck scatter_conv"; do
  set -- $spec
  timeout 280 bash tools/sanitize.sh $1 $2 > $O/${1}_$2.out 2>&1; echo "$1 $2 rc=$?"; tail -3 gpurun_out/sanitizer_${1}_$2.log | cut -c1-200
  cp gpurun_out/sanitizer_${1}_$2.log $O/ 2>/dev/null
done
