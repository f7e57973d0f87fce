#!/bin/bash
# final single-GPU validation of the tree: kernel checks, pytest -m gpu, smoke, headline + secondary benches
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only rmsnorm,zimage,layernorm > gpurun_out/check_rms.log 2>&1; cut -c1-200 gpurun_out/check_rms.log | tail -5
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-260 gpurun_out/bench_n1.json
timeout 600 python tools/bench_zimage.py --steps 4 --warmup 3 2> gpurun_out/zimage_ours_n1.err > gpurun_out/zimage_ours_n1.json; grep -o '{"metric.*' gpurun_out/zimage_ours_n1.json | cut -c1-260
timeout 900 python tools/bench_wan.py --steps 3 --warmup 3 2> gpurun_out/wan_ours_n1.err > gpurun_out/wan_ours_n1.json; grep -o '{"metric.*' gpurun_out/wan_ours_n1.json | cut -c1-260
