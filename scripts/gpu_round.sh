#!/bin/bash
# One GPU visit: parity tests, smoke, bench (both arms), kernel times, launch list.  Run under gpurun from the repo root.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
timeout 300 python scripts/kernel_times.py > gpurun_out/kernel_times.log 2>&1
timeout 300 python scripts/gemm_time.py > gpurun_out/gemm_times.log 2>&1
timeout 300 python scripts/conv_time.py > gpurun_out/conv_times.log 2>&1; tail -8 gpurun_out/conv_times.log
# launch list of the same step (eager launches so that every kernel is a separate ncu record; flush fills mark the step
# boundaries): only with GPU_ROUND_NCU=1 -- it takes ~2 minutes of box time
if [ "${GPU_ROUND_NCU:-0}" = "1" ]; then
ADAFACE_BENCH_GRAPH=0 ADAFACE_BENCH_EXTRAS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/launches.csv
fi
