"""Fixed vs per-unit cost of the short-context attention kernels: times the head-major cross path at growing sizes."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
H, S, d, C = 8, 77, 40, 320
def t(fn, flushing=True):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        if flushing: flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    return statistics.median(ts)
x = torch.zeros(64, device="cuda")
print("empty torch op (x.add_):", t(lambda: x.add_(1)), "us flushed;", t(lambda: x.add_(1), False), "us hot")
for B, N in ((1, 512), (1, 4096), (2, 4096), (4, 4096), (8, 4096), (16, 4096)):
    q = torch.zeros(B, H, N, 64, device="cuda", dtype=torch.bfloat16)
    q[..., :d] = torch.randn(B, H, N, d, device="cuda").to(torch.bfloat16)
    kv = torch.randn(B, S, 2 * C, device="cuda").to(torch.bfloat16)
    kh = kv[:, :, :C].unflatten(2, (H, d)).transpose(1, 2)
    vh = kv[:, :, C:].unflatten(2, (H, d)).transpose(1, 2)
    out = torch.empty(B, N, C, device="cuda", dtype=torch.bfloat16)
    f = lambda: a.ops.attention_headmajor(q, kh, vh, d ** -0.5, d=d, out=out)
    print(f"B={B} N={N}: flushed {t(f):.1f} us, hot {t(f, False):.1f} us, units={B*H*N//128}")
