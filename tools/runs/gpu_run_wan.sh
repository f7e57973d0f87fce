#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/bench_wan.py --steps 3 --warmup 3 > gpurun_out/wan_ours_n1.json 2> gpurun_out/wan_ours_n1.err; tail -c 1000 gpurun_out/wan_ours_n1.json; tail -3 gpurun_out/wan_ours_n1.err
timeout 1200 python tools/bench_wan.py --steps 3 --warmup 3 --impl reference > gpurun_out/wan_ref_n1.json 2> gpurun_out/wan_ref_n1.err; tail -c 1000 gpurun_out/wan_ref_n1.json; tail -3 gpurun_out/wan_ref_n1.err
