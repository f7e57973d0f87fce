#!/bin/bash
# one GPU call: kernel checks, timeline, pytest -m gpu, headline bench (bf16 + fp8)
mkdir -p gpurun_out
timeout 200 python tools/attn_trace.py 192 > gpurun_out/attn_trace.log 2>&1
timeout 300 python tools/gpu_check.py --only attention_speed > gpurun_out/attn_speed.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python bench.py --steps 5 --warmup 3 --dtype fp8 > gpurun_out/bench_n1_fp8.json 2> gpurun_out/bench_n1_fp8.err
cat gpurun_out/bench_n1.json gpurun_out/bench_n1_fp8.json | cut -c1-400
cat gpurun_out/attn_speed.log | cut -c1-600
