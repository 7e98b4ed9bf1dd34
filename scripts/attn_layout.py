"""Does the q/k/v memory layout bound the tcgen05 attention kernel?  interleaved [B,N,3C] vs head-major [3,B,H,N,d]."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return statistics.median(ts)
for N, C in ((4096, 320), (1024, 640)):
    B, H = 8, 8; d = C // H
    qkv = torch.randn(B, N, 3 * C, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    hm = torch.stack([t.reshape(B, N, H, d).permute(0, 2, 1, 3) for t in (q, k, v)]).contiguous()   # [3,B,H,N,d]
    dpad = (d + 63) // 64 * 64
    hp = torch.zeros(3, B, H, N, dpad, device="cuda", dtype=torch.bfloat16)
    hp[..., :d] = hm
    t3 = timeit(lambda: a.ops.attention_headmajor(hp[0], hp[1], hp[2], d ** -0.5, d=d))
    o3 = a.ops.attention_headmajor(hp[0], hp[1], hp[2], d ** -0.5, d=d)
    o1 = a.ops.attention(q, k, v, H, d ** -0.5)
    print("padded max diff", (o1.float() - o3.float()).abs().max().item(), f"padded head-major {t3*1e3:.1f} us ({4.0*B*N*N*C/t3/1e9:.0f} TF/s)")
    o2 = a.ops.attention_headmajor(hm[0], hm[1], hm[2], d ** -0.5)
    print("max diff", (o1.float() - o2.float()).abs().max().item())
    t1 = timeit(lambda: a.ops.attention(q, k, v, H, d ** -0.5))
    t2 = timeit(lambda: a.ops.attention_headmajor(hm[0], hm[1], hm[2], d ** -0.5))
    fl = 4.0 * B * N * N * C
    print(f"N={N} d={d}: interleaved {t1*1e3:.1f} us ({fl/t1/1e9:.0f} TF/s)   head-major {t2*1e3:.1f} us ({fl/t2/1e9:.0f} TF/s)", flush=True)
