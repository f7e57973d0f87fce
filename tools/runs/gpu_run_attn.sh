#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/attn_trace.py 192 > gpurun_out/attn_trace.log 2>&1
timeout 600 python tools/gpu_check.py > gpurun_out/check_all.log 2>&1
grep -E "attention|FAIL|checks passed" gpurun_out/check_all.log | cut -c1-700
head -3 gpurun_out/attn_trace.log | cut -c1-300
