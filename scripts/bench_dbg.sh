#!/bin/bash
# Step time of the headline workload with the GEMM epilogue stubbed (diagnosis: what the epilogue costs inside the graph)
for d in 0 1 2; do
ADAFACE_GEMM_DBG=$d ADAFACE_GEMM_BRES=0 ADAFACE_BENCH_DDIM=0 ADAFACE_BENCH_EXTRAS=0 ADAFACE_BENCH_STAGE2=0 timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('DBG=$d ms_per_step', d['ms_per_step'], 'headline', d['value'], 'e2e', d['e2e']['value'], 'check', d['check']['ok'])"
done
