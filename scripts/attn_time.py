"""Times the self-attention core at the four SD-1.5 levels (B=8) with L2 flushes; prints TFLOP/s."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for N, C in ((4096, 320), (1024, 640), (256, 1280), (64, 1280)):
    B, H = 8, 8
    qkv = torch.randn(B, N, 3 * C, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    for _ in range(3):
        a.ops.attention(q, k, v, H, (C // H) ** -0.5)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); a.ops.attention(q, k, v, H, (C // H) ** -0.5); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = statistics.median(ts)
    print(f"self-attn N={N} C={C} d={C//H}: {ms*1e3:.1f} us  {4.0*B*N*N*C/ms/1e9:.1f} TFLOP/s", flush=True)
# cross attention level A
N, C, S = 4096, 320, 77
q = torch.randn(8, N, C, device="cuda").to(torch.bfloat16)
kv = torch.randn(8, S, 2 * C, device="cuda").to(torch.bfloat16)
for _ in range(3):
    a.ops.attention(q, kv[:, :, :C], kv[:, :, C:], 8, 40 ** -0.5)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    flush.fill_(1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); a.ops.attention(q, kv[:, :, :C], kv[:, :, C:], 8, 40 ** -0.5); e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ms = statistics.median(ts)
print(f"cross-attn N={N} S={S}: {ms*1e3:.1f} us  {(2*8*N*C*2)/ms/1e6:.0f} GB/s (Q in + O out)")
