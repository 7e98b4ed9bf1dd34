#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "proj or gemm or lora or geglu" -p no:cacheprovider 2>&1 | grep -v Warn | tail -4
for b in 0 1; do echo "== ADAFACE_GEMM_BRES=$b"; ADAFACE_GEMM_BRES=$b timeout 200 python scripts/gemm_time.py 2>&1 | head -7; done
python - <<'PY'
import torch, statistics, sys, os
sys.path.insert(0, os.getcwd())
import adaface_dev_b200 as a
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
M, K = 32768, 320
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = torch.randn(2560, K, device="cuda").to(torch.bfloat16)
b = torch.zeros(2560, device="cuda")
for _ in range(3): a.ops.proj(x, w, bias=b, act=a.ops.ACT_GEGLU)
ts = []
for _ in range(8):
    flush.fill_(1); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); a.ops.proj(x, w, bias=b, act=a.ops.ACT_GEGLU); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ms = statistics.median(ts); print(f"A geglu M={M} N=2560 K={K}: {ms*1e3:.1f} us {2.0*M*2560*K/ms/1e9:.1f} TFLOP/s")
PY
ADAFACE_BENCH_DDIM=0 ADAFACE_BENCH_EXTRAS=0 ADAFACE_BENCH_STAGE2=0 timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('headline', d['value'], 'e2e', d['e2e']['value'], 'check', d['check'])"
