O=gpurun_out/b1; mkdir -p $O
timeout 300 python tools/profile_flux.py --fp8 --batch 1 --out $O/profile_flux_fp8_b1.json > $O/profile_flux_fp8_b1.txt 2>&1; head -30 $O/profile_flux_fp8_b1.txt
timeout 300 python bench.py --batch 1 --steps 10 --warmup 5 --no-bf16 > $O/bench_b1.json 2> $O/bench_b1.err; tail -c 200 $O/bench_b1.json
PA_DUAL_STREAM=0 timeout 300 python bench.py --batch 1 --steps 10 --warmup 5 --no-bf16 > $O/bench_b1_nodual.json 2> $O/bench_b1_nodual.err
python - <<'PY'
import json
for f in ("bench_b1","bench_b1_nodual"):
    d=json.load(open(f"gpurun_out/b1/{f}.json")); print(f, d["ms_per_step"], d["clocks"])
PY
