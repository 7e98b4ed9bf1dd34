#!/bin/bash
# One GPU visit: parity tests, smoke, bench (both arms).  Run under gpurun from the repo root.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v Warn | tail -80 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -8 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; head -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; head -c 600 gpurun_out/bench_ref.json
