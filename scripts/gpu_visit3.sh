#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_stage2.py tests/test_gpu_capture_consumers.py tests/test_gpu_ddim.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v Warning | tail -150 > gpurun_out/pytest_quick.log
tail -90 gpurun_out/pytest_quick.log
ADAFACE_BENCH_DDIM=0 ADAFACE_BENCH_EXTRAS=0 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_s2.json')); print(json.dumps(d.get('stage2_step'), indent=1))"; tail -5 gpurun_out/bench_s2.err
