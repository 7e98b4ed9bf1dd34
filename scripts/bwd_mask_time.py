"""Level-A self-attention backward (B=8, 4096 tokens, 8 x 40), unmasked vs key mask: us per call (delta + dk/dv + dq passes)."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
B, N, H, d = 8, 4096, 8, 40
C = H * d
qkv = torch.randn(B, N, 3 * C, device="cuda").to(torch.bfloat16)
do = torch.randn(B, N, C, device="cuda").to(torch.bfloat16)
q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
mask = (torch.rand(B, N, device="cuda") > 0.3).to(torch.uint8)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
dqkv = torch.empty_like(qkv)
for name, km in (("unmasked", None), ("key mask", mask)):
    lse = torch.empty(B, H, N, device="cuda")
    o = a.ops.attention(q, k, v, H, d ** -0.5, key_mask=km, lse=lse)
    f = lambda: a.ops.attention_bwd(q, k, v, o, do, lse, H, d ** -0.5, dqkv[:, :, :C], dqkv[:, :, C:2 * C], dqkv[:, :, 2 * C:], key_mask=km)
    for _ in range(2): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(6):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ms = statistics.median(ts)
    print(f"level-A self-attention backward {name:9s}: {ms*1e3:8.1f} us  {10.0*B*H*N*N*d/ms/1e9:7.1f} TFLOP/s (5 GEMMs)", flush=True)
