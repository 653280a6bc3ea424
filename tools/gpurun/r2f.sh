#!/bin/bash
# round 2, call F: ncu --set full of the tensor-bound GEMM shapes (stage-3/4 MLP) -- is the tensor pipe starved by the L2->SMEM fill?
cd /root/repo
for cfg in "24576 2048 512 gelu x bf16" "24576 512 2048 res x f32" "49152 1280 320 gelu x bf16" "49152 320 1280 res x f32"; do
  set -- $cfg
  name="r2f_gemm_${1}_${2}_${3}_${4}"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -f -o gpurun_out/$name python tools/gemm_one.py $1 $2 $3 $4 0 $6 > gpurun_out/$name.log 2>&1; echo "$name rc=$?"
done
ls -la gpurun_out/r2f_*.ncu-rep
