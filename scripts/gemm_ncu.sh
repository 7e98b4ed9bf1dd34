#!/bin/bash
mkdir -p gpurun_out
for b in 0 1; do
ADAFACE_GEMM_BRES=$b ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 1 -c 1 -o gpurun_out/gemm_bres$b -f python scripts/gemm_one.py > gpurun_out/ncu_gemm_bres$b.log 2>&1
tail -2 gpurun_out/ncu_gemm_bres$b.log
done
