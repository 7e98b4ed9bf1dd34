#!/bin/bash
# Short GPU visit for the U-Net convolutional blocks: a guarded first launch, the parity tests, then device times.
mkdir -p gpurun_out
timeout 120 python - > gpurun_out/conv_gate.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import adaface_dev_b200 as a
from oracle import unet_blocks_oracle as ub
for (B, h, w, ci, co, st) in ((1, 8, 16, 64, 64, 1), (2, 16, 16, 128, 64, 1), (2, 16, 16, 64, 64, 2)):
    x = torch.randn(B, ci, h, w).bfloat16().float(); wt = (torch.randn(co, ci, 3, 3) / (3 * ci ** 0.5)).bfloat16().float()
    y = a.ops.conv3x3(x.permute(0, 2, 3, 1).reshape(B, h * w, ci).contiguous().bfloat16().cuda(), a.ops.pack_conv3x3_weight(wt.cuda()), (h, w), stride=st, out_dtype=torch.float32)
    torch.cuda.synchronize()
    ref = ub.conv3x3(x, wt, None, stride=st)
    got = y.float().cpu().reshape(B, h // st, w // st, co).permute(0, 3, 1, 2)
    print((B, h, w, ci, co, st), "max err", (got - ref).abs().max().item(), flush=True)
PY
rc=$?
cat gpurun_out/conv_gate.log | tail -8
if [ $rc -ne 0 ]; then echo "GATE FAILED rc=$rc"; exit 0; fi
timeout 500 python -m pytest tests/test_gpu_unet_blocks.py -q --timeout 150 -p no:cacheprovider 2>&1 | tail -70 > gpurun_out/unet_pytest.log
tail -45 gpurun_out/unet_pytest.log
timeout 200 python scripts/conv_time.py > gpurun_out/conv_times.log 2>&1
cat gpurun_out/conv_times.log
