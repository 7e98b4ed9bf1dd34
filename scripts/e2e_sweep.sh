#!/bin/bash
run() { echo "== $*"; env "$@" ADAFACE_BENCH_DDIM=0 ADAFACE_BENCH_EXTRAS=0 ADAFACE_BENCH_STAGE2=0 timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('  device ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'e2e TFLOP/s', round(d['e2e']['value'],1))"; }
run X=1
run ADAFACE_BENCH_E2E_NOCOPY=1
run ADAFACE_BENCH_A_TAIL_SPLIT=8
run ADAFACE_BENCH_A_TAIL_SPLIT=2
