"""In-pipeline cost of the feed-forward / residual GEMMs of a BasicTransformerBlock (CUDA graph of 12 back-to-back launches, rotating
buffers): GEGLU up-projection (bias + exact-erf GELU gate) and the down-projection with bias + residual, levels A-C at B = 8."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
R, NB = 12, 3
def run(name, M, N, K, act, residual):
    xs = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in range(NB)]
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    b = torch.zeros(N, device="cuda")
    n_out = N // 2 if act == a.ops.ACT_GEGLU else N
    ys = [torch.empty(M, n_out, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
    rs = [torch.randn(M, n_out, device="cuda").to(torch.bfloat16) for _ in range(NB)] if residual else [None] * NB
    call = lambda i: a.ops.proj(xs[i % NB], w, bias=b, act=act, residual=rs[i % NB], out=ys[i % NB])
    for i in range(3): call(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(R): call(i)
    g.replay(); torch.cuda.synchronize(); ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) / R)
    ms = statistics.median(ts)
    print(f"{name:28s} M={M:6d} N={N:5d} K={K:5d}: {ms*1e3:7.1f} us/launch in a graph  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
for lvl, M, C in (("A", 32768, 320), ("B", 8192, 640), ("C", 2048, 1280)):
    run(f"{lvl} ff1 geglu (bias)", M, 8 * C, C, a.ops.ACT_GEGLU, False)
    run(f"{lvl} ff2 (bias + residual)", M, C, 4 * C, a.ops.ACT_NONE, True)
    run(f"{lvl} to_out (bias + residual)", M, C, C, a.ops.ACT_NONE, True)
