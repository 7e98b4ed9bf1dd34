"""Runs the level-A cross-attention fast path (B=8, 4096 queries, 77 keys, head-major padded q) a few times (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
B, N, C, H, S, d = 8, 4096, 320, 8, 77, 40
q = torch.zeros(B, H, N, 64, device="cuda", dtype=torch.bfloat16)
q[..., :d] = torch.randn(B, H, N, d, device="cuda").to(torch.bfloat16)
kv = torch.randn(B, S, 2 * C, device="cuda").to(torch.bfloat16)
kh = kv[:, :, :C].unflatten(2, (H, d)).transpose(1, 2)
vh = kv[:, :, C:].unflatten(2, (H, d)).transpose(1, 2)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    a.ops.attention_headmajor(q, kh, vh, d ** -0.5, d=d)
torch.cuda.synchronize()
