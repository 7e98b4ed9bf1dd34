"""SubjBasisGenerator forward at BASELINE config 2 (BS=64) as a CUDA graph, with programmatic dependent launch off / on."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
gen = a.SubjBasisGenerator().cuda().eval()
x = torch.randn(64, 16, 768, device="cuda") * 0.5
for pdl in (0, 1):
    a._lib.set_pdl(pdl)
    fn = a.graphed(lambda t: gen(t), x)
    for _ in range(3): fn(x)
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(x); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    print(f"SBG BS=64 graph, pdl={pdl}: {statistics.median(ts):.1f} us")
