#!/bin/bash
O=gpurun_out/ncuc; mkdir -p $O
timeout 120 ncu --set full --clock-control none --import-source on -c 1 --launch-skip 2 -k regex:xattn_cluster -o $O/ncu_xattn_cluster -f python tools/ncu_targets.py xattn > $O/ncu_xattn.log 2>&1; tail -1 $O/ncu_xattn.log
timeout 120 ncu --set full --clock-control none --import-source on -c 1 --launch-skip 2 -k regex:gn_cluster -o $O/ncu_gn_cluster -f python tools/ncu_targets.py groupnorm > $O/ncu_gn.log 2>&1; tail -1 $O/ncu_gn.log
