#!/bin/bash
# A/B of the GEMM / conv epilogue: LDS + STG.128 (ADAFACE_GEMM_TMA_STORE=0) vs TMA bulk stores (default).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet_blocks.py -x -q -m gpu 2>&1 | tail -5
for s in 0 1; do
  echo "== TMA_STORE=$s gemm"; ADAFACE_GEMM_TMA_STORE=$s timeout 300 python scripts/gemm_time.py 2>&1
  echo "== TMA_STORE=$s conv"; ADAFACE_GEMM_TMA_STORE=$s timeout 300 python scripts/conv_time.py 2>&1 | tail -20
done
