#!/bin/bash
# ncu --set full captures of the kernels added since run 17 + compute-sanitizer passes over their checks
mkdir -p gpurun_out
for t in gemm_2cta attn2 attn3 rmsmod; do
  timeout 400 ncu --set full --clock-control none --import-source on -c 1 --launch-skip 2 -k regex:'gemm_bf16_2cta_kernel|attention2_kernel|attention3_kernel|rms_mod_kernel' -o gpurun_out/ncu_$t -f python tools/ncu_targets.py $t > gpurun_out/ncu_$t.log 2>&1
  tail -1 gpurun_out/ncu_$t.log
done
for c in gemm_2cta attention_pair rmsnorm_modulate gemm_swiglu; do
  timeout 600 bash tools/sanitize.sh memcheck $c > gpurun_out/sanitize_memcheck_$c.out 2>&1; echo "memcheck $c rc=$?"
done
timeout 600 bash tools/sanitize.sh synccheck attention_pair > gpurun_out/sanitize_synccheck_attention_pair.out 2>&1; echo "synccheck attention_pair rc=$?"
timeout 600 bash tools/sanitize.sh racecheck rmsnorm_modulate > gpurun_out/sanitize_racecheck_rms.out 2>&1; echo "racecheck rms rc=$?"
ls gpurun_out/*.ncu-rep | tail -5
