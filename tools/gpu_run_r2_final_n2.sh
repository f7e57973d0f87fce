#!/bin/bash
# final 2-GPU validation of the tree: full GPU test suite + batch-1 sequence-parallel timing through the node API
O=gpurun_out/r2fin; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_gpu_2gpus.log; cat $O/pytest_gpu_2gpus.log
timeout 300 python bench.py --gpus 2 --api nodes --batch 1 --steps 8 --warmup 4 --no-bf16 > $O/nodes_ulysses_b1_fp8.json 2> $O/nodes_ulysses_b1_fp8.err
timeout 300 python bench.py --gpus 2 --api nodes --batch 1 --steps 8 --warmup 4 --dtype bf16 > $O/nodes_ulysses_b1_bf16.json 2> $O/nodes_ulysses_b1_bf16.err
for f in nodes_ulysses_b1_fp8 nodes_ulysses_b1_bf16; do echo "== $f"; tail -c 600 $O/$f.json; tail -c 300 $O/$f.err; done
