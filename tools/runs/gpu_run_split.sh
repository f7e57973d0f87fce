#!/bin/bash
O=gpurun_out/split; mkdir -p $O
timeout 240 python tools/mx8_pair_check.py > $O/mx8_pair_check.log 2>&1; rc=$?; tail -22 $O/mx8_pair_check.log; echo rc=$rc
[ $rc -ne 0 ] && exit 0
cp gpurun_out/mx8_pair_check.json $O/
timeout 300 python tools/gpu_check.py --only mxfp8_fused_quant_epilogues,flux_executor_fp8,gemm_mxfp8,wan_zimage_executors_fp8 --out $O/selfcheck.json > $O/selfcheck.log 2>&1; tail -2 $O/selfcheck.log | cut -c1-300
for sp in 1 0; do
PA_MXFP8_SPLITN=$sp PA_TIME=1 timeout 100 python tools/ncu_targets.py mxfp8_l1 2>&1 | grep us/launch
PA_MXFP8_SPLITN=$sp timeout 300 python bench.py --steps 6 --warmup 4 --no-bf16 > $O/bench_split$sp.json 2> $O/bench_split$sp.err
python -c "
import json; d=json.load(open('$O/bench_split$sp.json')); print('split=$sp', d['ms_per_step'], d['clocks']['sm_mhz'])"
done
