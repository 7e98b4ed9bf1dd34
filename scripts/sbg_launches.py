"""One SubjBasisGenerator forward at BS=64 (for an ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
gen = a.SubjBasisGenerator().cuda().eval()
x = torch.randn(64, 16, 768, device="cuda") * 0.5
with torch.no_grad():
    for _ in range(2): gen(x)
torch.cuda.synchronize()
