#!/bin/bash
O=gpurun_out/attproj; mkdir -p $O
timeout 300 python tools/gpu_check.py --only gemm_mxfp8_row_range_operand,flux_executor_fp8,gemm_mxfp8,mxfp8_fused_quant_epilogues --out $O/selfcheck.json > $O/selfcheck.log 2>&1; tail -6 $O/selfcheck.log | cut -c1-260
for v in 1 0; do
PA_FP8_ATT_PROJ=$v timeout 300 python bench.py --steps 6 --warmup 4 --no-bf16 > $O/bench_b8_att$v.json 2> $O/bench_b8_att$v.err
PA_FP8_ATT_PROJ=$v timeout 300 python bench.py --batch 1 --steps 10 --warmup 5 --no-bf16 > $O/bench_b1_att$v.json 2> $O/bench_b1_att$v.err
done
python - <<'PY'
import json
for f in ("bench_b8_att1","bench_b8_att0","bench_b1_att1","bench_b1_att0"):
    try:
        d=json.load(open(f"gpurun_out/attproj/{f}.json")); print(f, d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("gpu_launches"))
    except Exception as e: print(f, "ERR", e)
PY
