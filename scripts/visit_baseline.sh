#!/bin/bash
# Full GPU suite + smoke + bench + GEMM epilogue diagnosis (DBG: 1 = no global stores, 2 = no epilogue at all)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('headline', d['value'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], 'unet', d['unet_steps_per_s']['value'], 's2', d['stage2_step']['ms_per_iteration'])"
for b in 0 1; do for d in 0 1 2; do echo "== BRES=$b DBG=$d"; ADAFACE_GEMM_BRES=$b ADAFACE_GEMM_DBG=$d timeout 200 python scripts/gemm_time.py 2>&1 | head -4; done; done 2>&1 | tee gpurun_out/gemm_dbg.log
