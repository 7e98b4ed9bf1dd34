#!/bin/bash
# Short final visit (GPU budget nearly spent): full GPU suite, smoke, bench (our arm only).
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 600 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v Warn | tail -30 > gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; head -c 300 gpurun_out/bench.json; echo; tail -2 gpurun_out/bench.err
