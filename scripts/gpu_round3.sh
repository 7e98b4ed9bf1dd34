#!/bin/bash
# Round-2 final evidence visit: full GPU suite, smoke, bench (both arms), ncu launch list of the headline step, ncu --set full capture of
# the dominant kernel.  Run under gpurun from the repo root.
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v Warn | tail -40 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; head -c 400 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; head -c 300 gpurun_out/bench_ref.json; echo
ADAFACE_BENCH_DDIM=0 ADAFACE_BENCH_EXTRAS=0 ADAFACE_BENCH_STAGE2=0 ADAFACE_BENCH_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/launches_r02.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tcgen05_quad -s 2 -c 2 -o gpurun_out/r02_attn_quad -f python scripts/attn_one.py > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
timeout 100 python scripts/gemm_graph_time.py > gpurun_out/gemm_graph_times.log 2>&1
timeout 100 python scripts/attn_mask_time.py > gpurun_out/attn_mask_times.log 2>&1
timeout 300 python scripts/conv_time.py > gpurun_out/conv_times.log 2>&1; tail -4 gpurun_out/conv_times.log
