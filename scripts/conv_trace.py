"""Diagnosis: one traced launch of the implicit-GEMM 3x3 convolution (ADAFACE_GEMM_TRACE=1: CTA 0's hand-off stamps on stderr)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
B, side, cin, cout = 8, 64, 320, 320
x = torch.randn(B, side * side, cin, device="cuda").to(torch.bfloat16)
wp = a.ops.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device="cuda") * (9 * cin) ** -0.5)
bias = torch.zeros(cout, device="cuda")
os.environ.pop("ADAFACE_GEMM_TRACE", None)
for _ in range(3): a.ops.conv3x3(x, wp, (side, side), bias=bias)
torch.cuda.synchronize()
os.environ["ADAFACE_GEMM_TRACE"] = "1"
a.ops.conv3x3(x, wp, (side, side), bias=bias)
torch.cuda.synchronize()
