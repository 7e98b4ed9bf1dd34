"""Runs the capture cross-attention kernels at BASELINE config 1 (B=2, 4096 queries, 77 keys, fp32 q/k/v) a few times
(for ncu captures): forward with prob + score, then the backward with dprob."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
B, N, C, H, S = 2, 4096, 320, 8, 77
q, k, v = (torch.randn(B, n, C, device="cuda") for n in (N, S, S))
do, dprob = torch.randn(B, N, C, device="cuda").to(torch.bfloat16), torch.randn(B, H, N, S, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    a.ops.attention_cross_capture(q, k, v, H, 40 ** -0.5)
    a.ops.attention_cross_capture_bwd(q, k, v, do, H, 40 ** -0.5, dprob=dprob)
torch.cuda.synchronize()
