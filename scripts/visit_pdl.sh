#!/bin/bash
for lib in "" build/libadaface_late.so; do
echo "== lib=${lib:-default (early trigger)}"
ADAFACE_B200_LIB=$lib ADAFACE_BENCH_DDIM=0 ADAFACE_BENCH_EXTRAS=0 ADAFACE_BENCH_STAGE2=0 timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'headline', d['value'], 'e2e', d['e2e']['value'], 'check', d['check']['ok'])"
done
timeout 1200 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
