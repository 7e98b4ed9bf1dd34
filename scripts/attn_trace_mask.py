"""Diagnosis: one traced launch of the level-A self-attention kernel WITH a key mask (ADAFACE_ATTN_TRACE=1, -DAF_ATTN_TRACE build)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
N, C, B, H = 4096, 320, 8, 8
qkv = torch.randn(B, N, 3 * C, device="cuda").to(torch.bfloat16)
q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
km = (torch.rand(B, N, device="cuda") > 0.3).to(torch.uint8)
for _ in range(3): a.ops.attention(q, k, v, H, (C // H) ** -0.5, key_mask=km)
torch.cuda.synchronize()
os.environ["ADAFACE_ATTN_TRACE"] = "1"
a.ops.attention(q, k, v, H, (C // H) ** -0.5, key_mask=km)
torch.cuda.synchronize()
