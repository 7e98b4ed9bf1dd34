"""q / k|v projections + 77-key cross-attention at level A (B=8) as ONE CUDA graph of 20 repetitions: us per block, with the q projection
scattered head-major (ADAFACE_CROSS_HEADMAJOR=1, default) or written as plain [B, N, C] rows (=0)."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
torch.manual_seed(0)
for (N, C) in ((4096, 320), (1024, 640)):
    B, S, H = 8, 77, 8
    x = torch.randn(B * N, C, device="cuda").to(torch.bfloat16)
    ctx = torch.randn(B * S, 768, device="cuda").to(torch.bfloat16)
    wq = (torch.randn(C, C, device="cuda") * C ** -0.5).to(torch.bfloat16)
    wkv = (torch.randn(2 * C, 768, device="cuda") * 768 ** -0.5).to(torch.bfloat16)
    bq = torch.zeros(C, device="cuda"); bkv = torch.zeros(2 * C, device="cuda")
    REP = 20
    def body(x_, c_):
        o = None
        for _ in range(REP):
            o = a.ops.cross_attention_fused(x_, wq, bq, c_, wkv, bkv, B, N, S, H, (C // H) ** -0.5)
        return o
    fn = a.graphed(body, x, ctx)
    for _ in range(3): fn(x, ctx)
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); o = fn(x, ctx); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3 / REP)
    print(f"cross block N={N} C={C}: {statistics.median(ts):.1f} us per (q-proj + kv-proj + attention), checksum {o.float().abs().mean().item():.5f}", flush=True)
