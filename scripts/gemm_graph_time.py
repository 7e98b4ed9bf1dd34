"""In-pipeline cost of the projection GEMM: a CUDA graph of 20 back-to-back launches (rotating over 4 input / output buffer sets so
every launch reads data another launch of the graph wrote recently, as in the step graph), time per launch."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
shapes = [("A qkv", 32768, 960, 320), ("A out/q", 32768, 320, 320), ("B qkv", 8192, 1920, 640), ("B out/q", 8192, 640, 640),
          ("C qkv", 2048, 3840, 1280), ("C out/q", 2048, 1280, 1280), ("kv ctx A", 616, 640, 768)]
R, NB = 20, 4
for name, M, N, K in shapes:
    xs = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in range(NB)]
    w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    b = torch.zeros(N, device="cuda")
    ys = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
    for bias in (None, b):
        for i in range(3): a.ops.proj(xs[i % NB], w, bias=bias, out=ys[i % NB])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(R): a.ops.proj(xs[i % NB], w, bias=bias, out=ys[i % NB])
        g.replay(); torch.cuda.synchronize(); ts = []
        for _ in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); g.replay(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) / R)
        ms = statistics.median(ts)
        print(f"{name:9s} M={M:6d} N={N:5d} K={K:5d} bias={'y' if bias is not None else 'n'}: {ms*1e3:7.1f} us/launch in a graph  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
