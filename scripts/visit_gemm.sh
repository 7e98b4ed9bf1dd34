#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet_blocks.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
timeout 300 python scripts/gemm_time.py 2>&1 | tee gpurun_out/gemm_times.log
timeout 300 python scripts/conv_time.py 2>&1 | tail -22 | tee gpurun_out/conv_times.log
ADAFACE_GEMM_TMA_STORE=0 timeout 300 python scripts/gemm_time.py 2>&1 | head -4
timeout 100 python scripts/gemm_trace.py 32768 960 320 2>&1 | grep "^tile" | grep -v "[0-9]\{11,\}"
ADAFACE_BENCH_DDIM=0 ADAFACE_BENCH_EXTRAS=0 ADAFACE_BENCH_STAGE2=0 timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'headline', d['value'], 'e2e', d['e2e']['value'], 'check', d['check'])"
