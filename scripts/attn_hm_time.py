"""Level-A self-attention core on the padded head-major workspace layout (what the processor uses): us and TFLOP/s."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
B, H, N, d = 8, 8, 4096, 40
ws = torch.zeros(3, B, H, N, 64, device="cuda", dtype=torch.bfloat16)
ws[..., :d] = torch.randn(3, B, H, N, d, device="cuda").to(torch.bfloat16)
f = lambda: a.ops.attention_headmajor(ws[0], ws[1], ws[2], d ** -0.5, d=d)
for _ in range(3): f()
torch.cuda.synchronize(); ts = []
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
    flush.fill_(1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ms = statistics.median(ts)
print(f"head-major self-attn B={B} N={N} d={d}: {ms*1e3:.1f} us  {4.0*B*H*N*N*d/ms/1e9:.1f} TFLOP/s")
