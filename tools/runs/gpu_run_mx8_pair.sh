mkdir -p gpurun_out
timeout 200 python tools/mx8_pair_check.py > gpurun_out/mx8_pair_check.log 2>&1; rc=$?; tail -25 gpurun_out/mx8_pair_check.log; echo "rc=$rc"
if [ $rc -eq 0 ]; then
  PA_MXFP8_2CTA=1 timeout 300 python bench.py --steps 6 --warmup 4 --no-bf16 > gpurun_out/bench_fp8_pair1.json 2> gpurun_out/bench_fp8_pair1.err; tail -c 400 gpurun_out/bench_fp8_pair1.json
  PA_MXFP8_2CTA=0 timeout 300 python bench.py --steps 6 --warmup 4 --no-bf16 > gpurun_out/bench_fp8_pair0.json 2> gpurun_out/bench_fp8_pair0.err; tail -c 400 gpurun_out/bench_fp8_pair0.json
fi
