mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py --only attention_speed > gpurun_out/check17.log 2>&1
timeout 120 python tools/ncu_targets.py gemm_sdxl > gpurun_out/gemm_sdxl17.log 2>&1
for t in attn2 mxfp8 scatter conv gemm_sdxl; do
  timeout 400 ncu --set full --clock-control none --import-source on -c 1 --launch-skip 2 -k regex:'attention2_kernel|gemm_mxfp8_kernel|scatter_patch_embed_kernel|gemm_bf16_tcgen05_kernel' -o gpurun_out/ncu_$t -f python tools/ncu_targets.py $t > gpurun_out/ncu_$t.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
