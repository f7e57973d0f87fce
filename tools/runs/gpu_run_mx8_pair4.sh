#!/bin/bash
# after the CTA-pair MXFP8 kernel + straight-line QKV epilogue: numerics, headline, WAN / Z-Image fp8
O=gpurun_out/mx8pair4; mkdir -p $O
timeout 400 python tools/gpu_check.py --only mxfp8_fused_quant_epilogues,flux_executor_fp8,gemm_mxfp8 --out $O/selfcheck.json > $O/selfcheck.log 2>&1; tail -2 $O/selfcheck.log | cut -c1-200
PA_TIME=1 timeout 100 python tools/ncu_targets.py mxfp8_l1 > $O/l1_time.txt 2>&1; tail -1 $O/l1_time.txt
timeout 400 python bench.py --steps 8 --warmup 4 > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json
timeout 400 python tools/bench_wan.py --dtype fp8 --steps 3 --warmup 3 > $O/wan_fp8.json 2> $O/wan_fp8.err; tail -c 300 $O/wan_fp8.json
timeout 400 python tools/bench_zimage.py --dtype fp8 --steps 4 --warmup 3 > $O/zimage_fp8.json 2> $O/zimage_fp8.err; tail -c 300 $O/zimage_fp8.json
