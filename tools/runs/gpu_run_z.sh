#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -q -k zimage > gpurun_out/pytest_z.log 2>&1; tail -3 gpurun_out/pytest_z.log
timeout 600 python tools/bench_zimage.py --steps 4 --warmup 3 > gpurun_out/zimage_ours_n1.json 2> gpurun_out/zimage_ours_n1.err; tail -c 1500 gpurun_out/zimage_ours_n1.json; tail -3 gpurun_out/zimage_ours_n1.err
timeout 900 python tools/bench_zimage.py --steps 4 --warmup 3 --impl reference > gpurun_out/zimage_ref_n1.json 2> gpurun_out/zimage_ref_n1.err; tail -c 1200 gpurun_out/zimage_ref_n1.json; tail -3 gpurun_out/zimage_ref_n1.err
