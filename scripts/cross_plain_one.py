"""Runs the level-A cross-attention fast path (B=8, 4096 queries, 77 keys, q as plain [B, N, C] rows) a few times (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
B, N, C, H, S, d = 8, 4096, 320, 8, 77, 40
q = torch.randn(B, N, C, device="cuda").to(torch.bfloat16)
kv = torch.randn(B, S, 2 * C, device="cuda").to(torch.bfloat16)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    a.ops.attention(q, kv[:, :, :C], kv[:, :, C:], H, d ** -0.5)
torch.cuda.synchronize()
