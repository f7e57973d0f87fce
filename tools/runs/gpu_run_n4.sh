#!/bin/bash
mkdir -p gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $L --master-port 29551 bench.py --gpus 4 --steps 8 --warmup 4 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -c 700 gpurun_out/bench_n4.json
timeout 600 $L --master-port 29552 tools/bench_sdxl.py --gpus 4 --config 5 --steps 6 --warmup 3 > gpurun_out/sdxl5_ours_n4.json 2> gpurun_out/sdxl5_ours_n4.err; tail -c 900 gpurun_out/sdxl5_ours_n4.json; tail -2 gpurun_out/sdxl5_ours_n4.err
timeout 900 $L --master-port 29553 tools/bench_sdxl.py --gpus 4 --config 5 --steps 6 --warmup 3 --impl reference > gpurun_out/sdxl5_ref_n4.json 2> gpurun_out/sdxl5_ref_n4.err; tail -c 900 gpurun_out/sdxl5_ref_n4.json; tail -2 gpurun_out/sdxl5_ref_n4.err
