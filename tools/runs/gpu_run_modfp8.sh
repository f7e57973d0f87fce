#!/bin/bash
O=gpurun_out/modfp8; mkdir -p $O
timeout 400 python tools/gpu_check.py --only flux_executor_fp8,flux_executor --out $O/selfcheck.json > $O/selfcheck.log 2>&1; tail -3 $O/selfcheck.log | cut -c1-400
for m in 1 0; do
PA_FP8_MOD=$m timeout 300 python bench.py --batch 1 --steps 10 --warmup 5 --no-bf16 > $O/bench_b1_mod$m.json 2> $O/bench_b1_mod$m.err
PA_FP8_MOD=$m timeout 300 python bench.py --steps 6 --warmup 4 --no-bf16 > $O/bench_b8_mod$m.json 2> $O/bench_b8_mod$m.err
done
python - <<'PY'
import json
for f in ("bench_b1_mod1","bench_b1_mod0","bench_b8_mod1","bench_b8_mod0"):
    try:
        d=json.load(open(f"gpurun_out/modfp8/{f}.json")); print(f, d["ms_per_step"], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
PY
