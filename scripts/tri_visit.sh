#!/bin/bash
# A/B timings of the three-tile attention kernel; each line one process (the switches are read once).
run() { echo "== $*"; env "$@" ADAFACE_ATTN_TRI=1 timeout 100 python scripts/tri_check.py time 2>&1 | tail -1; }
IFS=';' read -ra CFGS <<< "${TRI_CFGS:-X=1;ADAFACE_TRI_N2=4;ADAFACE_TRI_N2=1;TRI_PDL=0;TRI_NOFLUSH=1}"
for cfg in "${CFGS[@]}"; do run $cfg; done
