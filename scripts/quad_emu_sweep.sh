#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention or attn or quad or lse" -p no:cacheprovider 2>&1 | grep -v Warn | tail -3
for emu in 2 3 5 6 7; do
  echo "== ADAFACE_EXP_EMU=$emu"
  ADAFACE_EXP_EMU=$emu timeout 120 python scripts/attn_time.py 2>&1 | head -1
done
