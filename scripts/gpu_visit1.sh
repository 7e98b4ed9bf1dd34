#!/bin/bash
# round-2 visit 1: parity tests (incl. DDIM + full-size U-Net), smoke, bench with the config-4 object, U-Net batch sweep
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -8 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | head -c 6000; tail -5 gpurun_out/bench.err
timeout 600 python scripts/unet_batch_sweep.py 8 16 32 64 > gpurun_out/unet_sweep.log 2>&1; tail -6 gpurun_out/unet_sweep.log
