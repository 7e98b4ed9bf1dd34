"""Runs the level-A flash backward (B=8, 4096 tokens, d=40, fused-QKV views) a few times (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
B, N, C, H = 8, 4096, 320, 8
d = C // H
qkv = torch.randn(B, N, 3 * C, device="cuda").to(torch.bfloat16)
do = torch.randn(B, N, C, device="cuda").to(torch.bfloat16)
q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
lse = torch.empty(B, H, N, device="cuda")
o = torch.empty(B, N, C, device="cuda", dtype=torch.bfloat16)
a.ops.attention(q, k, v, H, d ** -0.5, out=o, lse=lse)
dqkv = torch.empty_like(qkv)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    a.ops.attention_bwd(q, k, v, o, do, lse, H, d ** -0.5, dqkv[:, :, :C], dqkv[:, :, C:2 * C], dqkv[:, :, 2 * C:])
torch.cuda.synchronize()
