#!/bin/bash
# straight-line QKV epilogue: numerics (fused-quant + executor checks), ncu of linear1, step time
O=gpurun_out/mx8pair3; mkdir -p $O
timeout 400 python tools/gpu_check.py --only mxfp8_fused_quant_epilogues,flux_executor_fp8,gemm_mxfp8,wan_zimage_executors_fp8 --out $O/selfcheck.json > $O/selfcheck.log 2>&1; tail -8 $O/selfcheck.log
timeout 300 ncu --set full --clock-control none --import-source on -c 1 --launch-skip 2 -k regex:gemm_mxfp8 \
     -o $O/ncu_mxfp8_l1 -f python tools/ncu_targets.py mxfp8_l1 > $O/ncu_mxfp8_l1.log 2>&1; tail -1 $O/ncu_mxfp8_l1.log
timeout 300 python bench.py --steps 6 --warmup 4 --no-bf16 > $O/bench_fp8.json 2> $O/bench_fp8.err; tail -c 300 $O/bench_fp8.json
timeout 300 python tools/profile_flux.py --fp8 --out $O/profile_flux_fp8_b8.json > $O/profile_flux_fp8_b8.txt 2>&1; head -12 $O/profile_flux_fp8_b8.txt
