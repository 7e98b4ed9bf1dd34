"""Masked vs unmasked level-B self-attention (B=8, 1024 tokens, 8 x 80), us per call with L2 flushed."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
B, N, H, d = 8, 1024, 8, 80
C = H * d
qkv = torch.randn(B, N, 3 * C, device="cuda").to(torch.bfloat16)
q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
mask = (torch.rand(B, N, device="cuda") > 0.3).to(torch.uint8)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = torch.empty(B, N, C, device="cuda", dtype=torch.bfloat16)
for name, km in (("unmasked", None), ("key mask", mask)):
    for _ in range(3): a.ops.attention(q, k, v, H, d ** -0.5, key_mask=km, out=out)
    torch.cuda.synchronize(); ts = []
    for _ in range(8):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); a.ops.attention(q, k, v, H, d ** -0.5, key_mask=km, out=out); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ms = statistics.median(ts)
    print(f"level-B self-attention {name:9s}: {ms*1e3:7.1f} us  {4.0*B*H*N*N*d/ms/1e9:7.1f} TFLOP/s", flush=True)
