"""Per-kernel device times (CUDA events, L2 flushed between iterations) of the HBM-bound and backward kernels at
BASELINE sizes; prints microseconds and the algorithmic GB/s or TFLOP/s.  Usage: python scripts/kernel_times.py [filter]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a

ops = a.ops
BF = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
only = sys.argv[1] if len(sys.argv) > 1 else ""


def timeit(name, fn, bytes_=None, flops=None, iters=10):
    if only and only not in name:
        return
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = statistics.median(ts)
    extra = ""
    if bytes_:
        extra += f"  {bytes_ / ms / 1e6:8.0f} GB/s"
    if flops:
        extra += f"  {flops / ms / 1e9:8.1f} TFLOP/s"
    print(f"{name:58s} {ms * 1e3:9.1f} us{extra}", flush=True)


def rn(*s, dt=BF):
    return torch.randn(*s, device="cuda").to(dt)


H, S = 8, 77
# ---- cross-attention, level A (HBM-bound): fast path (bf16) and capture path (fp32 q/k/v, prob + score out)
for B in (2, 8):
    N, C = 4096, 320
    d = C // H
    q, kv = rn(B, N, C), rn(B, S, 2 * C)
    core = 2 * B * N * C * 2 + 2 * B * S * C * 2
    timeit(f"cross fast (interleaved q)   B={B} N={N} d={d}", lambda: ops.attention(q, kv[:, :, :C], kv[:, :, C:], H, d ** -0.5), bytes_=core)
    ws = torch.zeros(B, H, N, 64, device="cuda", dtype=BF)
    kh = kv[:, :, :C].unflatten(2, (H, d)).transpose(1, 2)
    vh = kv[:, :, C:].unflatten(2, (H, d)).transpose(1, 2)
    timeit(f"cross fast (head-major q)    B={B} N={N} d={d}", lambda: ops.attention_headmajor(ws, kh, vh, d ** -0.5, d=d),
           bytes_=B * N * H * 64 * 2 + B * N * C * 2)
    timeit(f"cross stream bf16 (interleaved)  B={B} N={N} d={d}",
           lambda: ops.attention_cross_capture(q, kv[:, :, :C], kv[:, :, C:], H, d ** -0.5, want_prob=False, want_score=False), bytes_=core)
    qf, kf, vf = rn(B, N, C, dt=torch.float32), rn(B, S, C, dt=torch.float32), rn(B, S, C, dt=torch.float32)
    maps = B * H * N * S * 4
    cap_core = B * N * C * 4 + B * N * C * 2 + 2 * B * S * C * 4
    timeit(f"capture fwd fp32 prob+score  B={B} N={N}", lambda: ops.attention_cross_capture(qf, kf, vf, H, d ** -0.5), bytes_=cap_core + 2 * maps)
    timeit(f"capture fwd fp32 prob only   B={B} N={N}", lambda: ops.attention_cross_capture(qf, kf, vf, H, d ** -0.5, want_score=False),
           bytes_=cap_core + maps)
    timeit(f"capture fwd fp32 no maps     B={B} N={N}", lambda: ops.attention_cross_capture(qf, kf, vf, H, d ** -0.5, want_prob=False, want_score=False),
           bytes_=cap_core)
    do, dprob = rn(B, N, C), rn(B, H, N, S, dt=torch.float32)
    bwd_core = B * N * C * (4 + 2 + 2) + 2 * B * S * C * 4
    timeit(f"capture bwd fp32 dprob       B={B} N={N}", lambda: ops.attention_cross_capture_bwd(qf, kf, vf, do, H, d ** -0.5, dprob=dprob), bytes_=bwd_core + maps)
    timeit(f"capture bwd fp32 dO only     B={B} N={N}", lambda: ops.attention_cross_capture_bwd(qf, kf, vf, do, H, d ** -0.5), bytes_=bwd_core)
    for t_ in (ops.chan_major,):
        timeit(f"chan_major fp32 [B,N,C]->[B,C,N] B={B}", lambda: ops.chan_major(qf, 0.5), bytes_=2 * B * N * C * 4)

# ---- flash attention forward with lse + backward (self-attention core)
for B, N, C in ((1, 4096, 320), (8, 4096, 320), (8, 1024, 640), (8, 256, 1280)):
    d = C // H
    qkv, do = rn(B, N, 3 * C), rn(B, N, C)
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    lse = torch.empty(B, H, N, device="cuda")
    o = torch.empty(B, N, C, device="cuda", dtype=BF)
    fl = 4.0 * B * N * N * C
    timeit(f"flash fwd +lse  B={B} N={N} d={d}", lambda: ops.attention(q, k, v, H, d ** -0.5, out=o, lse=lse), flops=fl)
    dqkv = torch.empty_like(qkv)
    timeit(f"flash bwd       B={B} N={N} d={d}",
           lambda: ops.attention_bwd(q, k, v, o, do, lse, H, d ** -0.5, dqkv[:, :, :C], dqkv[:, :, C:2 * C], dqkv[:, :, 2 * C:]), flops=2.5 * fl)

# ---- helpers
x = rn(32768, 320)
w, b = torch.ones(320, device="cuda"), torch.zeros(320, device="cuda")
timeit("layernorm fwd bf16 [32768,320]", lambda: ops.layernorm(x, w, b), bytes_=2 * x.numel() * 2)
timeit("layernorm bwd bf16 [32768,320]", lambda: ops.layernorm_bwd(x, x, w), bytes_=3 * x.numel() * 2)
timeit("transpose bf16 [32768,320]", lambda: ops.transpose(x), bytes_=2 * x.numel() * 2)
timeit("colsum bf16 [32768,320]", lambda: ops.colsum(x), bytes_=x.numel() * 2)
u = rn(32768, 2560)
timeit("geglu act fwd [32768,2560]->[32768,1280]", lambda: ops.act_fwd(u, ops.ACT_GEGLU), bytes_=u.numel() * 2 * 1.5)
