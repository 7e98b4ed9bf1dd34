#!/bin/bash
# run a subset of the GPU tests with full failure detail:  bash scripts/gpu_quick_tests.sh "<pytest -k expression or paths>"
mkdir -p gpurun_out
timeout 1500 python -m pytest $1 -q -m gpu --timeout 600 -p no:cacheprovider -x 2>&1 | grep -v Warning | tail -150 > gpurun_out/pytest_quick.log
tail -120 gpurun_out/pytest_quick.log
