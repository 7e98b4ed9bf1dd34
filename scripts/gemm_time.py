"""Times the projection GEMM at the SD-1.5 attention-stack shapes (B=8); prints TFLOP/s and GB/s."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
shapes = [("A qkv", 32768, 960, 320), ("A out/q", 32768, 320, 320), ("B qkv", 8192, 1920, 640), ("B out/q", 8192, 640, 640),
          ("C qkv", 2048, 3840, 1280), ("C out/q", 2048, 1280, 1280), ("kv ctx A", 616, 640, 768), ("kv ctx C", 616, 2560, 768),
          ("sbg qkv", 1280, 2304, 768), ("sbg fc1", 1280, 3072, 768), ("sbg fc2", 1280, 768, 3072)]
for name, M, N, K in shapes:
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    for _ in range(3): a.ops.proj(x, w)
    torch.cuda.synchronize(); ts = []
    for _ in range(8):
        if not os.environ.get("NOFLUSH"): flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); a.ops.proj(x, w); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ms = statistics.median(ts)
    print(f"{name:9s} M={M:6d} N={N:5d} K={K:5d}: {ms*1e3:7.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s  {(M*K+N*K+M*N)*2/ms/1e6:7.0f} GB/s", flush=True)
