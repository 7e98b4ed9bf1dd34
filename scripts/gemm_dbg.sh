#!/bin/bash
for b in 0 1; do for d in 0 1 2; do echo "== BRES=$b DBG=$d"; ADAFACE_GEMM_BRES=$b ADAFACE_GEMM_DBG=$d timeout 200 python scripts/gemm_time.py 2>&1 | head -2; done; done
