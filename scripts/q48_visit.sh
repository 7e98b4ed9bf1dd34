#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention or attn or quad or lse" -p no:cacheprovider 2>&1 | grep -v Warn | tail -5
IFS=';' read -ra CFGS <<< "${Q48_CFGS:-Q48=0;EMU=0 DEG=2;EMU=2 DEG=2;EMU=4 DEG=2}"
for cfg in "${CFGS[@]}"; do
  q48=1; emu=3; deg=2
  for kv in $cfg; do case $kv in Q48=*) q48=${kv#Q48=};; EMU=*) emu=${kv#EMU=};; DEG=*) deg=${kv#DEG=};; esac; done
  echo "== $cfg"
  ADAFACE_ATTN_Q48=$q48 ADAFACE_EXP_EMU=$emu ADAFACE_EXP_DEG=$deg timeout 120 python scripts/attn_time.py 2>&1 | head -1
done
