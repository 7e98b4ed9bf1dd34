"""Diagnosis: one traced launch of the projection GEMM (ADAFACE_GEMM_TRACE=1 makes the library dump CTA 0's hand-off stamps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
M, N, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (32768, 960, 320)))
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
os.environ.pop("ADAFACE_GEMM_TRACE", None)
for _ in range(3): a.ops.proj(x, w)
if not os.environ.get("NOFLUSH"): flush.fill_(1)
torch.cuda.synchronize()
os.environ["ADAFACE_GEMM_TRACE"] = "1"
a.ops.proj(x, w)
torch.cuda.synchronize()
