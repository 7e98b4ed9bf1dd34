#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -8 gpurun_out/smoke.log
ADAFACE_BENCH_DDIM=0 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); s=d.get('secondary',{}); print({k:(v.get('us'),v.get('frac')) for k,v in s.items() if isinstance(v,dict) and 'us' in v}); print(d['value'], d['e2e']['value'])"; tail -5 gpurun_out/bench.err
