#!/bin/bash
O=gpurun_out/sp2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_engine.py -x -q -k "ulysses" 2>&1 | tail -15 > $O/pytest_ulysses.log; cat $O/pytest_ulysses.log
